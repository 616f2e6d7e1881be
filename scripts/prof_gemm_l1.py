"""Three launches of the layer-1 forward transform of cfg2 (30,926 x 256 -> 256, ReLU + dropout) for ncu."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from primekg_rgcn_linkprediction_b200 import ops

DEV = "cuda:0"
n, K, N = 30926, 256, 256
P = ops.alloc_planes(n, K, "fp32", DEV)
ops.split_planes(torch.randn(n, K, device=DEV), P)
W = torch.randn(K, N, device=DEV) * 0.05
bias = torch.zeros(N, device=DEV)
wp = ops.prepare_weights(W, None, "fp32")
out = torch.empty(n, N, device=DEV)
ctr = ops.dropout_counter(DEV)
for _ in range(3):
    ops.transform_fwd_w(P, K, wp, N, bias, True, "fp32", 0.5, 1, ctr, out=out)
torch.cuda.synchronize()
print("ok")
