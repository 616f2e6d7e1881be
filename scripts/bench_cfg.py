"""Step time (forward + BCE + backward, CUDA-graph replay) for the non-headline configs of BASELINE.json:
cfg1 (64->128->128), cfg2, cfg3 (full-PrimeKG-shaped, 30 relations, num_bases = 8)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import primekg_rgcn_linkprediction_b200 as pkg
from primekg_rgcn_linkprediction_b200 import synth

DEV = "cuda:0"
which = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
mode = sys.argv[2] if len(sys.argv) > 2 else "fp32"
if which == "cfg1":
    kg, hid, bases = synth.primekg_subgraph(), 128, None
elif which == "cfg2":
    kg, hid, bases = synth.primekg_subgraph(), 256, None
else:
    kg, hid, bases = synth.primekg_full(), 256, 8
heads, tails, rels, labels = synth.link_batch(kg, 1024)
torch.manual_seed(42)
model = pkg.DrugDiseaseModel(kg.num_nodes, kg.num_relations, 64, hid, dropout=0.5, decoder_dropout=0.1, num_bases=bases).to(DEV)
for c in (model.encoder.conv1, model.encoder.conv2):
    c.mode = mode
model.train()
ei, et = kg.edge_index.to(DEV), kg.edge_type.to(DEV)
step = pkg.GraphedTrainStep(model, ei, et, batch_size=2048)
step.load_batch(heads.to(DEV), tails.to(DEV), rels.to(DEV), labels.to(DEV))
flush = torch.empty(128 * 1024 * 1024, dtype=torch.float32, device=DEV)
for _ in range(3):
    step()
torch.cuda.synchronize()
ts = []
for _ in range(10):
    flush.fill_(0.0)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); step(); b.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
ms = sum(ts) / len(ts)
print(json.dumps({"config": which, "mode": mode, "nodes": kg.num_nodes, "edges": kg.num_edges, "relations": kg.num_relations,
                  "hidden": hid, "num_bases": bases, "ms_per_step": ms, "edges_per_sec": kg.num_edges / (ms * 1e-3),
                  "loss": float(step.loss), "max_mem_GB": torch.cuda.max_memory_allocated() / 1e9}))
