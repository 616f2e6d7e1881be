"""Warm per-kernel timeline of one CUDA-graph replay of the cfg2 step (torch.profiler / CUPTI): durations and the idle
gaps between consecutive kernels.  Writes profiles-style text to stdout."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

import primekg_rgcn_linkprediction_b200 as pkg
from primekg_rgcn_linkprediction_b200 import synth

DEV = "cuda:0"
which = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
hid = 128 if which == "cfg1" else 256
kg = synth.primekg_subgraph()
heads, tails, rels, labels = (t.to(DEV) for t in synth.link_batch(kg, 1024))
torch.manual_seed(42)
model = pkg.DrugDiseaseModel(kg.num_nodes, kg.num_relations, 64, hid, dropout=0.5, decoder_dropout=0.1).to(DEV)
model.train()
ei, et = kg.edge_index.to(DEV), kg.edge_type.to(DEV)
step = pkg.GraphedTrainStep(model, ei, et, batch_size=2048)
step.load_batch(heads, tails, rels, labels)
for _ in range(5):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range is not None]
evs.sort(key=lambda e: e.time_range.start)
# keep the last replay: split on large gaps
if not evs:
    print("no CUDA events recorded")
    sys.exit(0)
groups, cur = [], [evs[0]]
for a, b in zip(evs, evs[1:]):
    if b.time_range.start - a.time_range.end > 100:     # us: between replays
        groups.append(cur); cur = []
    cur.append(b)
groups.append(cur)
g = groups[-1]
t0 = g[0].time_range.start
busy = sum(e.time_range.end - e.time_range.start for e in g)
span = g[-1].time_range.end - t0
print(f"# {which}: one graph replay, {len(g)} kernels, span {span:.1f} us, busy {busy:.1f} us, idle {span - busy:.1f} us")
prev_end = t0
for e in g:
    d = e.time_range.end - e.time_range.start
    gap = e.time_range.start - prev_end
    print(f"{e.time_range.start - t0:9.1f} us  dur {d:7.1f}  gap {gap:5.1f}  {e.name[:90]}")
    prev_end = e.time_range.end
