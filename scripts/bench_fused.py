"""Fused layer forward (schedule 3) against walk -> transform (schedule 1) on the cfg2 layer shapes; CUDA events,
back-to-back launches after warm-up.  python scripts/bench_fused.py [fp32|bf16]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import primekg_rgcn_linkprediction_b200 as pkg                      # noqa: E402
from primekg_rgcn_linkprediction_b200 import ops, synth            # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "fp32"
which = sys.argv[2] if len(sys.argv) > 2 else "cfg2"
dev = "cuda"
kg = synth.primekg_subgraph() if which == "cfg2" else synth.primekg_full()
g = pkg.RelGraph.from_edges(kg.edge_index.to(dev), kg.edge_type.to(dev), kg.num_nodes, kg.num_relations)
R = kg.num_relations
torch.manual_seed(0)
for d_in, d_out in ((64, 256), (256, 256)):
    x = torch.randn(kg.num_nodes, d_in, device=dev)
    W = torch.randn(R * d_in, d_out, device=dev) * 0.05
    root = torch.randn(d_in, d_out, device=dev) * 0.05
    bias = torch.randn(d_out, device=dev)
    ref = None
    for schedule in (1, 3):
        ctr = ops.dropout_counter(x.device)
        run = lambda: ops.layer_fwd(g, x, x, W, root, bias, True, mode, 0.5, 7, ctr, pipeline=schedule)
        for _ in range(5):
            out = run()[0]
        torch.cuda.synchronize()
        n = 50
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            for _ in range(n):
                out = run()[0]
        gr.replay(); torch.cuda.synchronize()
        e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1000 / n
        print(f"{which} {mode} {d_in}->{d_out} schedule {schedule}: {us:.1f} us per layer forward", flush=True)
