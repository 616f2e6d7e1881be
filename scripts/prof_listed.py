"""Three listed-rows layer forwards of cfg2's layer 2 (for ncu: -k regex:aggregate_rows_kernel --launch-skip 1 -c 1)."""
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
N, R, E = kg.num_nodes, kg.num_relations, kg.num_edges
gen = torch.Generator().manual_seed(0)
x = torch.randn(N, 256, generator=gen).to(DEV)
W = (torch.randn(R * 256, 256, generator=gen) / 16).to(DEV)
root = (torch.randn(256, 256, generator=gen) / 16).to(DEV)
bias = torch.zeros(256, device=DEV)
pos = torch.randint(0, E, (1024,), generator=gen)
head = torch.cat([kg.edge_index[0, pos], kg.edge_index[0, pos]]).to(DEV)
tail = torch.cat([kg.edge_index[1, pos], torch.randint(0, N, (1024,), generator=gen)]).to(DEV)
rows, slot = ops.rows_list_build(head, tail, N)
for _ in range(3):
    ops.layer_fwd(g, x, x, W, root, bias, False, "fp32", rows=rows, slot=slot)
torch.cuda.synchronize()
print("ok")
