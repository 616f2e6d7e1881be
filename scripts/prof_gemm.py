"""Small driver for ncu: the three cfg2 layer-2 tensor-core kernels, a few launches each."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from primekg_rgcn_linkprediction_b200 import ops

DEV = "cuda:0"
mode = sys.argv[1] if len(sys.argv) > 1 else "fp32"
torch.manual_seed(0)
n, K1, K2, N = 30926, 768, 256, 256
A = ops.alloc_planes(n, K1 + K2, mode, DEV)
ops.split_planes(torch.randn(n, K1 + K2, device=DEV), A)
G = ops.alloc_planes(n, N, mode, DEV)
part = ops.split_planes(torch.randn(n, N, device=DEV), G, colsum=True)
W1 = torch.randn(K1, N, device=DEV)
W2 = torch.randn(K2, N, device=DEV)
b = torch.randn(N, device=DEV)
for _ in range(3):
    ops.transform_fwd(A, K1, K2, W1, W2, b, True, mode)
    ops.transform_dgrad(G, N, W1, W2, mode)
    ops.transform_wgrad(A, K1, K2, G, N, part, mode)
torch.cuda.synchronize()
print("done")
