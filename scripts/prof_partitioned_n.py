"""Kernel-time table of one partitioned step on N GPUs (torchrun; rank 0 prints): the cfg5-shaped graph sized per GPU
(1.25 M nodes / 50 M edges per rank, 30 relations, 3 layers), peer-memory exchange.  Run with RGCN_PDL=0 so that a kernel's
duration does not include the wait for its predecessor."""
import collections
import os
import re
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

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
nodes, edges, R, L, B = 1_250_000 * world, 50_000_000 * world, 30, 3, 2048
kg = synth.scaled_kg(nodes, edges, R, seed=42, device=dev)
plan = D.plan_partition(kg.edge_index[1], nodes, world)
src, dst, rel = D.local_edges(kg.edge_index, kg.edge_type, plan, rank)
del kg
torch.cuda.empty_cache()
graph = pkg.RelGraph(src, dst, rel, plan.max_n, world * plan.max_n, R)
del src, dst, rel
model = DF.FusedPartitionedModel(plan, rank, R, 64, 128, dropout=0.0, decoder_dropout=0.0, num_layers=L, seed=42).to(dev)
model.encoder.set_graph(graph)
model.train()
g = torch.Generator(device=dev).manual_seed(7)
heads = torch.randint(0, nodes, (B,), generator=g, device=dev)
tails = torch.randint(0, nodes, (B,), generator=g, device=dev)
rels = torch.randint(0, R, (B,), generator=g, device=dev)
labels = (torch.rand(B, generator=g, device=dev) < 0.5).float()
sl = slice(rank * B // world, (rank + 1) * B // world)


def step():
    for p in model.parameters():
        p.grad = None
    sc = model(heads[sl], tails[sl], rels[sl])
    F.binary_cross_entropy_with_logits(sc, labels[sl], reduction="sum").div(B).backward()
    model.allreduce_decoder_grads()


for _ in range(2):
    step()
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); step(); b.record(); torch.cuda.synchronize()
step_ms = a.elapsed_time(b)
dist.barrier()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
dist.barrier()
# every rank: one line with its own kernel time by category (is some rank systematically slower than the others?)
cat = collections.defaultdict(float)
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        n = e.name
        k = ("barrier" if "barrier_kernel" in n else "walk_fwd" if ("aggregate_rows" in n and "false, false, false" in n.split("<")[-1] and ", 0, false" in n)
             else "walk_bwd" if "aggregate_rows" in n else "hub" if "hub_partial" in n else "gemm_fwd" if "gemm_kmajor_kernel<true, true" in n
             else "dgrad" if "gemm_kmajor" in n else "wgrad" if "wgrad" in n else "pull" if ("reduce_split" in n or "pull_rows" in n)
             else "push" if "push_rows" in n else "other")
        cat[k] += e.device_time / 1e3
line = f"rank {rank}: step {step_ms:7.2f} ms | " + " ".join(f"{k} {v:6.2f}" for k, v in sorted(cat.items()))
lines = [None] * world
dist.all_gather_object(lines, line)
if rank == 0:
    print("\n".join(lines))
if rank == 0:
    tot = collections.defaultdict(lambda: [0.0, 0])
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA:
            name = re.sub(r"\(.*", "", e.name).replace("rgcn::", "").replace("void ", "")[:70]
            tot[name][0] += e.device_time
            tot[name][1] += 1
    total = sum(v[0] for v in tot.values())
    print(f"# {world} GPUs, rank 0, one step: {step_ms:.2f} ms by events, {total / 1e3:.2f} ms of kernel time")
    for name, (t, n) in sorted(tot.items(), key=lambda kv: -kv[1][0])[:28]:
        print(f"{t / 1e3:9.3f} ms  {100 * t / total:5.1f} %  x{n:<3d} {name}")
dist.destroy_process_group()
