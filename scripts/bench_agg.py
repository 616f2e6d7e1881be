"""Cold-L2 timing of the aggregation kernels on the cfg2 graph (d = 64 and 256, forward and backward)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import primekg_rgcn_linkprediction_b200 as pkg
from primekg_rgcn_linkprediction_b200 import ops, synth

DEV = "cuda:0"
kg = synth.primekg_subgraph()
ei, et = kg.edge_index.to(DEV), kg.edge_type.to(DEV)
g = pkg.get_graph(ei, et, kg.num_nodes, kg.num_relations)
print("hubs fwd", g.fwd.n_hubs, g.fwd.n_chunks, "bwd", g.bwd.n_hubs, g.bwd.n_chunks, "max seg", g.max_seg, g.max_seg_t)
flush = torch.empty(128 * 1024 * 1024, dtype=torch.float32, device=DEV)
E, N, R = kg.num_edges, kg.num_nodes, kg.num_relations
for d in (64, 256):
    x = torch.randn(N, d, device=DEV)
    gA = torch.randn(N, (R + 1) * d, device=DEV)
    A = ops.alloc_planes(N, (R + 1) * d, "fp32", DEV)
    for name, fn, nbytes in (("fwd fp32-out", lambda: ops.aggregate_fwd(g, x), E * (d * 4 + 4) + (N * R + 1) * 4),
                             ("fwd planes", lambda: ops.aggregate_fwd(g, x, planes=A), E * (d * 4 + 4) + (N * R + 1) * 4),
                             ("bwd", lambda: ops.aggregate_bwd(g, gA, d, init=gA[:, R * d:]), E * (d * 4 + 8) + (N * R + 1) * 4)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        # a short kernel finishes before the Python host has enqueued the closing event (one ops call costs ~40 us of
        # host time): capture REP calls into a CUDA graph and time the replay, so the number is the GPU's
        REP = 8
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                for _ in range(REP):
                    fn()
        torch.cuda.current_stream().wait_stream(side)
        ts = []
        for _ in range(10):
            flush.fill_(0.0)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); gr.replay(); b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) / REP)
        ms = sum(ts) / len(ts)
        print(f"d={d} {name}: {ms*1e3:.1f} us (hub + rows, graph replay of {REP}, L2 cold for the first)  {nbytes/ms/1e6:.0f} GB/s algorithmic", flush=True)
