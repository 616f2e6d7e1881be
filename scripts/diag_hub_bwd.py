"""Why is the backward hub pass of the cfg5 shard ~50x slower than the forward one?  Times hub pass + row walk, forward and
backward, on the 1.25 M-node / 50 M-edge / 30-relation shard graph and prints the hub plans."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import primekg_rgcn_linkprediction_b200 as pkg
from primekg_rgcn_linkprediction_b200 import ops, synth

DEV = "cuda:0"
n, e, R, d = 1_250_000, 50_000_000, 30, 128
kg = synth.scaled_kg(n, e, R, seed=42, device=DEV)
g = pkg.RelGraph.from_edges(kg.edge_index, kg.edge_type, n, R)
del kg
print("fwd hubs", g.fwd.n_hubs, "chunks", g.fwd.n_chunks, "| bwd hubs", g.bwd.n_hubs, "chunks", g.bwd.n_chunks, flush=True)
x = torch.randn(n, d, device=DEV)


def t(fn, it=3):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it):
        fn()
    b.record(); b.synchronize()
    return a.elapsed_time(b) / it


print("aggregate_fwd d=128 (bf16 out)", round(t(lambda: ops.aggregate_fwd(g, x, out_bf16=True)), 3), "ms", flush=True)
gA = torch.randn(n, (R + 1) * d, device=DEV)
print("aggregate_bwd d=128, gA [n, 31 d]", round(t(lambda: ops.aggregate_bwd(g, gA, d, init=gA[:, R * d:])), 3), "ms", flush=True)
del gA
torch.cuda.empty_cache()
# the same walk over a NARROW matrix (relation stride 0 is not expressible; use R = 1 columns): gather table 0.64 GB
gB = torch.randn(n, d, device=DEV)
gsmall = pkg.RelGraph.from_edges(torch.stack([g.col.long()[:1000], g.col.long()[:1000]]), torch.zeros(1000, dtype=torch.long, device=DEV), n, 1)
del gsmall
print("done")
