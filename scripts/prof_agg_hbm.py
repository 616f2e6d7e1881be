"""Small driver for ncu: the DRAM-bound gather of bench.py's roofline.hbm_bound_case (cfg5-shard shape: 4 M nodes /
64 M edges / 30 relations, uniform sources, d = 128 => 2 GB of features), forward walk + the bare gather probe."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import primekg_rgcn_linkprediction_b200 as pkg
from primekg_rgcn_linkprediction_b200 import ops, synth

DEV = "cuda:0"
n, e, r, d = 4_000_000, 64_000_000, 30, 128
big = synth.scaled_kg(n, e, r, seed=7, power=1.0, device=DEV)
g = pkg.RelGraph.from_edges(big.edge_index, big.edge_type, n, r)
del big
x = torch.randn(n, d, device=DEV)
for _ in range(2):
    ops.aggregate_fwd(g, x, out_bf16=True)
    ops.probe_gather(x, g.col, blocks_per_sm=8)
torch.cuda.synchronize()
print("done")
