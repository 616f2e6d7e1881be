"""Small driver for ncu: the cfg2 last-layer backward, row-sparse form (compaction, dgrad / wgrad over 4,096 rows, the
transposed walk that skips absent edges) and dense form, a few launches each."""
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
N, R, d = kg.num_nodes, kg.num_relations, 256
heads, tails, _, _ = synth.link_batch(kg, 1024)
rows = torch.cat([heads, tails]).to(DEV)
x = torch.randn(N, d, device=DEV)
W = torch.randn(R * d, d, device=DEV) / 16
root = torch.randn(d, d, device=DEV) / 16
bias = torch.zeros(d, device=DEV)
_, A = ops.layer_fwd(g, x, x, W, root, bias, False, "fp32")
gO = torch.zeros(N, d, device=DEV)
gO[rows.unique()] = torch.randn(rows.unique().numel(), d, device=DEV)
os.environ["RGCN_OVERLAP_WGRAD"] = "0"
for _ in range(3):
    ops.layer_bwd(g, gO, None, 1.0, A, W, root, d, "fp32", True, True, True, True, rows=rows)
    ops.layer_bwd(g, gO, None, 1.0, A, W, root, d, "fp32", True, True, True, True)
torch.cuda.synchronize()
print("done")
