"""Diagnostic (not a test): prints the error of each tcgen05 transform kernel per shape/mode, so that ONE
GPU run tells which operand path (K-major A, TMA B, MN-major wgrad, epilogue) is wrong if any."""
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


torch.manual_seed(0)
for (n, K1, K2, N) in [(128, 64, 0, 32), (128, 64, 0, 64), (300, 48, 16, 32), (1000, 192, 64, 128),
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
            out = ops.transform_fwd(A1, A2, W1, W2, b, False, mode)
            torch.cuda.synchronize()
            res["fwd"] = err(out, A.double() @ W.double() + b.double())
            gA = ops.transform_dgrad(gO, None, W1, W2, mode)
            torch.cuda.synchronize()
            res["dgrad"] = err(gA, gO.double() @ W.double().t())
            gW1, gW2, gb = ops.transform_wgrad(A1, A2, gO, None, mode)
            torch.cuda.synchronize()
            res["wgrad"] = err(gW1, A1.double().t() @ gO.double())
            if K2:
                res["wgrad2"] = err(gW2, A2.double().t() @ gO.double())
            res["gbias"] = err(gb, gO.double().sum(0))
        except Exception as ex:
            res["EXC"] = repr(ex)[:300]
        print((n, K1, K2, N), mode, {k: (f"{v:.2e}" if isinstance(v, float) else v) for k, v in res.items()}, flush=True)

# timing at cfg2 sizes (layer 1: 64 -> 256, layer 2: 256 -> 256), relu mask on the backward kernels as in training
for tag, (n, K1, K2, N) in (("cfg2-L1", (30926, 192, 64, 256)), ("cfg2-L2", (30926, 768, 256, 256))):
    A1 = torch.randn(n, K1, device=DEV)
    A2 = torch.randn(n, K2, device=DEV)
    W1 = torch.randn(K1, N, device=DEV)
    W2 = torch.randn(K2, N, device=DEV)
    b = torch.randn(N, device=DEV)
    gO = torch.randn(n, N, device=DEV)
    ro = torch.randn(n, N, device=DEV).clamp(min=0)
    flush = torch.empty(128 * 1024 * 1024, dtype=torch.float32, device=DEV)
    for mode in ("fp32", "bf16"):
        for name, fn in (("fwd", lambda: ops.transform_fwd(A1, A2, W1, W2, b, True, mode)),
                         ("dgrad", lambda: ops.transform_dgrad(gO, ro, W1, W2, mode)),
                         ("wgrad", lambda: ops.transform_wgrad(A1, A2, gO, ro, mode))):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            ts = []
            for _ in range(10):
                flush.fill_(0.0)
                a, bb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                fn()
                bb.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(bb))
            ms = sum(ts) / len(ts)
            print(f"{tag} {name} {mode}: {ms*1e3:.1f} us cold-L2 ({2*n*(K1+K2)*N/ms/1e9:.1f} TFLOP/s useful)", flush=True)
