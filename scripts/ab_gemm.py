"""A/B of the layer-forward transform: round-1 form (weights converted + transposed per call, K-major B) against the
prepared-weights form (MN-major B), cfg2 layer shapes, L2 flushed, CUDA events."""
import os
import sys
import statistics

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from primekg_rgcn_linkprediction_b200 import ops

DEV = "cuda:0"
flush_buf = torch.empty(128 * 1024 * 1024, dtype=torch.float32, device=DEV)


def t(fn, iters=30):
    for _ in range(5):
        fn()
    ts = []
    for _ in range(iters):
        flush_buf.fill_(1.0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return round(statistics.median(ts), 1)


for mode in ("fp32", "bf16"):
    for n, K, N in ((30926, 256, 256), (30926, 1024, 256), (30926, 512, 128), (4096, 1024, 256)):
        A = torch.randn(n, K, device=DEV)
        P = ops.alloc_planes(n, K, mode, DEV)
        ops.split_planes(A, P)
        W = torch.randn(K, N, device=DEV) * 0.05
        bias = torch.zeros(N, device=DEV)
        wp = ops.prepare_weights(W, None, mode)
        out = torch.empty(n, N, device=DEV)
        G = ops.alloc_planes(n, N, mode, DEV)
        ops.split_planes(torch.randn(n, N, device=DEV), G)
        print(mode, (n, K, N), "fwd old(conv+kmajor)", t(lambda: ops.transform_fwd(P, K, 0, W, None, bias, True, mode)),
              "prepare", t(lambda: ops.prepare_weights(W, None, mode)),
              "fwd_w(mn-major)", t(lambda: ops.transform_fwd_w(P, K, wp, N, bias, True, mode, out=out)),
              "dgrad old", t(lambda: ops.transform_dgrad(G, N, W, None, mode)),
              "dgrad_w", t(lambda: ops.transform_dgrad(G, N, W, None, mode, w_planes=wp)), "us", flush=True)
