"""Listed-rows forward of cfg2's layer 2 (30,926 x 256 features, 4,096 head / tail rows drawn like the bench's batches)
against the dense layer: walk and transform timed through CUDA-graph replay.  A/B the walk's grid split with
RGCN_LIST_SLICE (column slice width) and RGCN_LIST_RSPLIT (relation ranges per row), read once per process."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import primekg_rgcn_linkprediction_b200 as pkg
from primekg_rgcn_linkprediction_b200 import ops, synth

DEV = "cuda:0"
mode = sys.argv[1] if len(sys.argv) > 1 else "fp32"
kg = synth.primekg_subgraph()
ei, et = kg.edge_index.to(DEV), kg.edge_type.to(DEV)
g = pkg.get_graph(ei, et, kg.num_nodes, kg.num_relations)
N, R, E = kg.num_nodes, kg.num_relations, kg.num_edges
gen = torch.Generator().manual_seed(0)
d_in = d_out = 256
x = torch.randn(N, d_in, generator=gen).to(DEV)
W = (torch.randn(R * d_in, d_out, generator=gen) / 16).to(DEV)
root = (torch.randn(d_in, d_out, generator=gen) / 16).to(DEV)
bias = torch.zeros(d_out, device=DEV)
# positives = random edges (degree-proportional), negatives = corrupted tails, as bench.py's batches
pos = torch.randint(0, E, (1024,), generator=gen)
head = torch.cat([kg.edge_index[0, pos], kg.edge_index[0, pos]]).to(DEV)
tail = torch.cat([kg.edge_index[1, pos], torch.randint(0, N, (1024,), generator=gen)]).to(DEV)
rows, slot = ops.rows_list_build(head, tail, N)
deg = torch.bincount(kg.edge_index[1], minlength=N)
print("listed rows", rows.numel(), "unique", torch.unique(rows).numel(), "edges into listed rows",
      int(deg[torch.unique(rows.cpu())].sum()), "of", E)
x16 = ops.to_bf16(x) if mode == "bf16" else None
flush = torch.empty(128 * 1024 * 1024, dtype=torch.float32, device=DEV)


def timed(fn, name, rep=8):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            for _ in range(rep):
                fn()
    torch.cuda.current_stream().wait_stream(side)
    ts = []
    for _ in range(10):
        flush.fill_(0.0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); gr.replay(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) / rep)
    print(f"{name}: {sum(ts) / len(ts) * 1e3:.1f} us", flush=True)


print("RGCN_LIST_SLICE", os.environ.get("RGCN_LIST_SLICE"), "RGCN_LIST_RSPLIT", os.environ.get("RGCN_LIST_RSPLIT"), "mode", mode)
timed(lambda: ops.layer_fwd(g, x, x, W, root, bias, False, mode, x_bf16=x16), "dense layer forward ")
timed(lambda: ops.layer_fwd(g, x, x, W, root, bias, False, mode, x_bf16=x16, rows=rows, slot=slot), "listed layer forward")
out_d = ops.layer_fwd(g, x, x, W, root, bias, False, mode, x_bf16=x16)[0]
out_l = ops.layer_fwd(g, x, x, W, root, bias, False, mode, x_bf16=x16, rows=rows, slot=slot)[0]
print("listed rows bit-equal to dense:", bool(torch.equal(out_d[rows], out_l[rows])))
