"""Small driver for ncu: the cfg2 layer-1 aggregation (gather width 64: 256-byte rows), forward and backward."""
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
N, R, d = kg.num_nodes, kg.num_relations, 64
x = torch.randn(N, d, device=DEV)
gA = torch.randn(N, (R + 1) * d, device=DEV)
A = ops.alloc_planes(N, (R + 1) * d, "fp32", DEV)
for _ in range(3):
    ops.aggregate_fwd(g, x, planes=A)
    ops.aggregate_bwd(g, gA, d, init=gA[:, R * d:])
torch.cuda.synchronize()
print("done")
