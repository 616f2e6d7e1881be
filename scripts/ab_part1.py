"""Step time of the partitioned path on ONE GPU (cfg5 shard shape: 1.25 M nodes / 50 M edges / 30 relations / 3 layers),
CUDA events over 5 steps; for A/B runs of environment switches (RGCN_STREAM_REL, PRIMEKG_RGCN_SPARSE_FWD ...)."""
import os
import socket
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import torch.nn.functional as F

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
    loss = F.binary_cross_entropy_with_logits(sc, labels, reduction="sum").div(B)
    loss.backward()
    return loss


W = int(os.environ.get("AB_WARMUP", "2"))
for _ in range(W):
    step()
torch.cuda.synchronize()
per = []
for _ in range(int(os.environ.get("AB_STEPS", "5"))):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); loss = step(); b.record(); torch.cuda.synchronize()
    per.append(round(a.elapsed_time(b), 1))
print("per-step ms:", per)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    loss = step()
b.record(); torch.cuda.synchronize()
gn = float(sum(p.grad.double().square().sum() for p in model.parameters()).sqrt())
print(f"RGCN_STREAM_REL={os.environ.get('RGCN_STREAM_REL')} partitioned 1-GPU step: {a.elapsed_time(b) / 5:.2f} ms, loss {float(loss):.6f}, grad norm {gn:.6e}")
dist.destroy_process_group()
