"""Kernel-time table of one partitioned step on ONE GPU (world size 1: the per-GPU work of the cfg5 shard, no exchange):
which kernels the 1.25 M-node / 50 M-edge / 30-relation / 3-layer step spends its time in (torch.profiler / CUPTI)."""
import collections
import os
import re
import socket
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import torch.nn.functional as F
from torch.profiler import ProfilerActivity, profile

import primekg_rgcn_linkprediction_b200 as pkg
from primekg_rgcn_linkprediction_b200 import dist as D
from primekg_rgcn_linkprediction_b200 import dist_fused as DF
from primekg_rgcn_linkprediction_b200 import synth

dev = torch.device("cuda", 0)
s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
dist.init_process_group("nccl", rank=0, world_size=1, device_id=dev)
nodes, edges, R, L, B = 1_250_000, 50_000_000, 30, 3, 2048
kg = synth.scaled_kg(nodes, edges, R, seed=42, device=dev)
plan = D.plan_partition(kg.edge_index[1], nodes, 1)
src, dst, rel = D.local_edges(kg.edge_index, kg.edge_type, plan, 0)
del kg
graph = pkg.RelGraph(src, dst, rel, plan.max_n, plan.max_n, R)
del src, dst, rel
model = DF.FusedPartitionedModel(plan, 0, R, 64, 128, dropout=0.0, decoder_dropout=0.0, num_layers=L, seed=42).to(dev)
model.encoder.set_graph(graph)
model.train()
g = torch.Generator(device=dev).manual_seed(7)
heads = torch.randint(0, nodes, (B,), generator=g, device=dev)
tails = torch.randint(0, nodes, (B,), generator=g, device=dev)
rels = torch.randint(0, R, (B,), generator=g, device=dev)
labels = (torch.rand(B, generator=g, device=dev) < 0.5).float()


def step():
    for p in model.parameters():
        p.grad = None
    sc = model(heads, tails, rels)
    F.binary_cross_entropy_with_logits(sc, labels, reduction="sum").div(B).backward()


for _ in range(2):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
tot = collections.defaultdict(lambda: [0.0, 0])
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        name = re.sub(r"\(.*", "", e.name).replace("rgcn::", "").replace("void ", "")[:70]
        tot[name][0] += e.device_time
        tot[name][1] += 1
total = sum(v[0] for v in tot.values())
print(f"# one step: {total / 1e3:.2f} ms of kernel time")
for name, (t, n) in sorted(tot.items(), key=lambda kv: -kv[1][0])[:25]:
    print(f"{t / 1e3:9.3f} ms  {100 * t / total:5.1f} %  x{n:<3d} {name}")
dist.destroy_process_group()
