"""Host-side profile (cProfile) of the eager module-API step: where the Python time per step goes."""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

import primekg_rgcn_linkprediction_b200 as pkg
from primekg_rgcn_linkprediction_b200 import synth

DEV = "cuda:0"
kg = synth.primekg_subgraph()
heads, tails, rels, labels = (t.to(DEV) for t in synth.link_batch(kg, 1024))
torch.manual_seed(42)
model = pkg.DrugDiseaseModel(kg.num_nodes, kg.num_relations, 64, 256, dropout=0.5, decoder_dropout=0.1).to(DEV)
model.train()
ei, et = kg.edge_index.to(DEV), kg.edge_type.to(DEV)
params = list(model.parameters())


def step():
    for p in params:
        p.grad = None
    s = model(ei, et, heads, tails, rels)
    loss = F.binary_cross_entropy_with_logits(s, labels)
    loss.backward()
    return loss


for _ in range(20):
    step()
torch.cuda.synchronize()
n = 300
t0 = time.perf_counter()
for _ in range(n):
    step()
t_host = time.perf_counter() - t0
torch.cuda.synchronize()
t_all = time.perf_counter() - t0
print(f"host enqueue {t_host / n * 1e3:.3f} ms/step, wall incl. GPU {t_all / n * 1e3:.3f} ms/step")
pr = cProfile.Profile()
pr.enable()
for _ in range(n):
    step()
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(28)
