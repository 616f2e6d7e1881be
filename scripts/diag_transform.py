"""Diagnostic (not a test): prints the error and the cold-L2 time of each tcgen05 transform kernel per shape / mode."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import __graft_entry__ as e

e.build()
from primekg_rgcn_linkprediction_b200 import ops

DEV = "cuda:0"


def err(g, w):
    return float((g.double() - w).abs().max() / (w.abs().max() + 1e-30))


def planes(mats, mode, relu_mask=None, colsum=False):
    n = mats[0].size(0)
    P = ops.alloc_planes(n, sum(m.size(1) for m in mats), mode, mats[0].device)
    c0, part = 0, None
    for m in mats:
        part = ops.split_planes(m, P, col0=c0, relu_mask=relu_mask, colsum=colsum)
        c0 += m.size(1)
    return P, part


torch.manual_seed(0)
for (n, K1, K2, N) in [(128, 64, 0, 32), (128, 64, 0, 64), (300, 48, 16, 32), (100, 36, 12, 20), (1000, 192, 64, 128),
                       (4097, 768, 256, 256), (130, 64, 0, 512)]:
    for mode in ("bf16", "fp32"):
        A1 = torch.randn(n, K1, device=DEV)
        A2 = torch.randn(n, K2, device=DEV) if K2 else None
        W1 = torch.randn(K1, N, device=DEV) / (K1 + K2) ** .5
        W2 = torch.randn(K2, N, device=DEV) / (K1 + K2) ** .5 if K2 else None
        b = torch.randn(N, device=DEV)
        gO = torch.randn(n, N, device=DEV)
        W = torch.cat([W1, W2], 0) if K2 else W1
        A = torch.cat([A1, A2], 1) if K2 else A1
        res = {}
        try:
            Ap, _ = planes([A1] + ([A2] if K2 else []), mode)
            Gp, part = planes([gO], mode, colsum=True)
            out = ops.transform_fwd(Ap, K1, K2, W1, W2, b, False, mode)
            torch.cuda.synchronize()
            res["fwd"] = err(out, A.double() @ W.double() + b.double())
            gA = ops.transform_dgrad(Gp, N, W1, W2, mode)
            torch.cuda.synchronize()
            res["dgrad"] = err(gA, gO.double() @ W.double().t())
            gW1, gW2, gb = ops.transform_wgrad(Ap, K1, K2, Gp, N, part, mode)
            torch.cuda.synchronize()
            res["wgrad"] = err(gW1, A1.double().t() @ gO.double())
            if K2:
                res["wgrad2"] = err(gW2, A2.double().t() @ gO.double())
            res["gbias"] = err(gb, gO.double().sum(0))
        except Exception as ex:
            res["EXC"] = repr(ex)[:300]
        print((n, K1, K2, N), mode, {k: (f"{v:.2e}" if isinstance(v, float) else v) for k, v in res.items()}, flush=True)

# timing at cfg2 sizes (layer 1: 64 -> 256, layer 2: 256 -> 256)
flush = torch.empty(128 * 1024 * 1024, dtype=torch.float32, device=DEV)


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.fill_(0.0)
        a, bb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        bb.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(bb))
    return sum(ts) / len(ts)


for tag, (n, K1, K2, N) in (("cfg2-L1", (30926, 192, 64, 256)), ("cfg2-L2", (30926, 768, 256, 256))):
    A1 = torch.randn(n, K1, device=DEV)
    A2 = torch.randn(n, K2, device=DEV)
    W1 = torch.randn(K1, N, device=DEV)
    W2 = torch.randn(K2, N, device=DEV)
    b = torch.randn(N, device=DEV)
    gO = torch.randn(n, N, device=DEV)
    ro = torch.randn(n, N, device=DEV).clamp(min=0)
    for mode in ("fp32", "bf16"):
        Ap, _ = planes([A1, A2], mode)
        Gp, part = planes([gO], mode, relu_mask=ro, colsum=True)
        G2 = ops.alloc_planes(n, N, mode, DEV)
        for name, fn in (("split_planes(gO,mask)", lambda: ops.split_planes(gO, G2, relu_mask=ro, colsum=True)),
                         ("fwd", lambda: ops.transform_fwd(Ap, K1, K2, W1, W2, b, True, mode)),
                         ("dgrad", lambda: ops.transform_dgrad(Gp, N, W1, W2, mode)),
                         ("wgrad", lambda: ops.transform_wgrad(Ap, K1, K2, Gp, N, part, mode))):
            ms = timeit(fn)
            print(f"{tag} {name} {mode}: {ms*1e3:.1f} us cold-L2 ({2*n*(K1+K2)*N/ms/1e9:.1f} TFLOP/s useful)", flush=True)
