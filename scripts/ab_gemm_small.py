"""GPU-bound timing (CUDA-graph replay of 8 launches, L2 flushed before the first) of the transforms that sit on the
critical path of the cfg2 step: layer-1 forward (30,926 x 256 -> 256, ReLU + dropout), layer-1 dgrad, the listed
forward (4,096 x 1,024 -> 256) and the compact dgrad (4,096 x 256 -> 1,024)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from primekg_rgcn_linkprediction_b200 import ops

DEV = "cuda:0"
flush = torch.empty(128 * 1024 * 1024, dtype=torch.float32, device=DEV)


def timed(fn, name, rep=8):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            for _ in range(rep):
                fn()
    torch.cuda.current_stream().wait_stream(side)
    ts = []
    for _ in range(10):
        flush.fill_(0.0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); gr.replay(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) / rep)
    print(f"{name}: {sorted(ts)[len(ts) // 2] * 1e3:.1f} us", flush=True)


mode = sys.argv[1] if len(sys.argv) > 1 else "fp32"
ctr = ops.dropout_counter(DEV)
for n, K, N in ((30926, 256, 256), (4096, 1024, 256), (30926, 1024, 256)):
    P = ops.alloc_planes(n, K, mode, DEV)
    ops.split_planes(torch.randn(n, K, device=DEV), P)
    W = torch.randn(K, N, device=DEV) * 0.05
    bias = torch.zeros(N, device=DEV)
    wp = ops.prepare_weights(W, None, mode)
    out = torch.empty(n, N, device=DEV)
    G = ops.alloc_planes(n, N, mode, DEV)
    ops.split_planes(torch.randn(n, N, device=DEV), G)
    timed(lambda: ops.transform_fwd_w(P, K, wp, N, bias, True, mode, out=out), f"{mode} {(n, K, N)} fwd_w relu")
    timed(lambda: ops.transform_fwd_w(P, K, wp, N, bias, True, mode, 0.5, 1, ctr, out=out), f"{mode} {(n, K, N)} fwd_w relu+dropout")
    timed(lambda: ops.transform_dgrad(G, N, W, None, mode, w_planes=wp), f"{mode} {(n, K, N)} dgrad_w (n x {N} -> {K})")
