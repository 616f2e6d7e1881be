"""BASELINE cfg4: score all 6,282 x 5,593 drug-disease pairs (DistMult and cosine) and the ranking evaluation of
reference src/evaluate.py:219-291 (15,372 test edges against all 30,926 entities) on the B200, GPU-bound timing through
CUDA-graph replay, with the CPU oracle timed beside it."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import primekg_rgcn_linkprediction_b200 as pkg
from oracle import rgcn_ref as O

DEV = "cuda:0"
d = int(sys.argv[1]) if len(sys.argv) > 1 else 256
torch.manual_seed(3)
emb = torch.randn(30926, d, device=DEV)
drugs = torch.arange(5593, 11875, device=DEV)
diseases = torch.arange(0, 5593, device=DEV)
rel = torch.randn(d, device=DEV)
table = torch.randn(3, d, device=DEV)
nq = 15372
heads = torch.randint(0, 30926, (nq,), device=DEV)
tails = torch.randint(0, 30926, (nq,), device=DEV)
rels = torch.zeros(nq, dtype=torch.int64, device=DEV)


def gpu_time(fn, rep=5, iters=5):
    fn(); torch.cuda.synchronize()
    side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            for _ in range(rep):
                fn()
    torch.cuda.current_stream().wait_stream(side)
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); gr.replay(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) / rep)
    return min(ts)


pairs = drugs.numel() * diseases.numel()
out = {"d": d, "pairs": pairs}
for name, fn in (("distmult_all_pairs", lambda: pkg.score_all_pairs(emb, drugs, diseases, rel_vec=rel)),
                 ("cosine_all_pairs", lambda: pkg.score_all_pairs(emb, drugs, diseases, cosine=True))):
    ms = gpu_time(fn)
    out[name] = {"ms": round(ms, 4), "pairs_per_s": pairs / (ms * 1e-3), "tflops": 2 * pairs * d / (ms * 1e-3) / 1e12}
ms = gpu_time(lambda: pkg.rank_true_tails(emb, table, heads, rels, tails), rep=2, iters=3)
out["rank_15372_queries_vs_30926"] = {"ms": round(ms, 4), "pairs_per_s": nq * 30926 / (ms * 1e-3),
                                        "tflops": 2 * nq * 30926 * d / (ms * 1e-3) / 1e12}
# CPU oracle beside it (all host cores): the [6282, 5593] score matrix, and the reference's rank loop on a sample of rows
ce, cdr, cdi, crel = emb.cpu(), drugs.cpu(), diseases.cpu(), rel.cpu()
t0 = time.perf_counter(); O.distmult_allpairs_ref(ce, cdr, cdi, crel); t = time.perf_counter() - t0
out["cpu_distmult_all_pairs"] = {"ms": round(t * 1e3, 2), "pairs_per_s": pairs / t, "cores": os.cpu_count()}
ns = 1024
h = ce[heads[:ns].cpu()] * table.cpu()[0]
t0 = time.perf_counter()
S = h @ ce.t()
for i in range(ns):
    torch.argsort(S[i], descending=True)
t = time.perf_counter() - t0
out["cpu_rank_loop_sample"] = {"queries": ns, "ms": round(t * 1e3, 1), "pairs_per_s": ns * 30926 / t,
                               "note": "score_all_tails + per-row argsort as at src/evaluate.py:260-276"}
print(json.dumps(out))
