"""Aggregation timing on a UNIFORM graph of cfg2's size (no hubs, no long rows): separates the per-edge throughput of the
row walk from the effect of the degree skew."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import primekg_rgcn_linkprediction_b200 as pkg
from primekg_rgcn_linkprediction_b200 import ops, synth

DEV = "cuda:0"
for name, kg in (("uniform", synth.uniform_kg(30926, 849456, 3, seed=1)), ("primekg", synth.primekg_subgraph())):
    ei, et = kg.edge_index.to(DEV), kg.edge_type.to(DEV)
    g = pkg.get_graph(ei, et, kg.num_nodes, kg.num_relations)
    print(name, "hubs", g.fwd.n_hubs, g.fwd.n_chunks, "max seg", g.max_seg)
    flush = torch.empty(128 * 1024 * 1024, dtype=torch.float32, device=DEV)
    E, N, R = kg.num_edges, kg.num_nodes, kg.num_relations
    for d in (64, 256):
        x = torch.randn(N, d, device=DEV)
        gA = torch.randn(N, (R + 1) * d, device=DEV)
        A = ops.alloc_planes(N, (R + 1) * d, "fp32", DEV)
        for nm, fn, nbytes in (("fwd planes", lambda: ops.aggregate_fwd(g, x, planes=A), E * (d * 4 + 4) + (N * R + 1) * 4),
                               ("bwd", lambda: ops.aggregate_bwd(g, gA, d, init=gA[:, R * d:]), E * (d * 4 + 8) + (N * R + 1) * 4)):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            ts = []
            for _ in range(10):
                flush.fill_(0.0)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); fn(); b.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            ms = sum(ts) / len(ts)
            print(f"  {name} d={d} {nm}: {ms*1e3:.1f} us  {nbytes/ms/1e6:.0f} GB/s algorithmic", flush=True)
