"""Dense relational transform of one RGCN layer (the `h_r @ W_r` products + root term).

    forward  O  = H @ Wf + X @ root + bias            H = [H_0 | ... | H_{R-1}]   [N, R*d_in]
    dgrad    gA = gO @ [Wf ; root]^T                  [N, (R+1)*d_in]
    wgrad    gWf = H^T @ gO ,  g_root = X^T @ gO ,  g_bias = sum_i gO[i]

Replaces the R+1 matmuls per layer of RGCNConv's loop path (reference call sites
src/models/rgcn.py:123, :128) and their autograd transposes (src/train.py:306).

mode "fp32": fp32 in, fp32 accumulate.   mode "bf16": operands rounded to bf16, fp32 accumulate
(the "bf16-transform" mode of BASELINE.json, tolerance 2e-2).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

MODES = ("fp32", "bf16")


def _mm(a: torch.Tensor, b: torch.Tensor, mode: str) -> torch.Tensor:
    if mode == "bf16":
        return torch.mm(a.to(torch.bfloat16), b.to(torch.bfloat16), out_dtype=torch.float32)
    return torch.mm(a, b)


def transform_fwd(H: torch.Tensor, Wf: torch.Tensor, x: torch.Tensor, root: torch.Tensor, bias: torch.Tensor,
                  relu: bool, mode: str) -> torch.Tensor:
    out = _mm(H, Wf, mode)
    out += _mm(x, root, mode)
    out += bias
    if relu:
        out.relu_()
    return out


def transform_dgrad(gO: torch.Tensor, Wcat: torch.Tensor, mode: str) -> torch.Tensor:
    """gO [N, d_out], Wcat [(R+1)*d_in, d_out] -> [N, (R+1)*d_in]."""
    return _mm(gO, Wcat.t(), mode)


def transform_wgrad(H: torch.Tensor, x: torch.Tensor, gO: torch.Tensor, mode: str
                    ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    gWf = _mm(H.t(), gO, mode)
    g_root = _mm(x.t(), gO, mode)
    return gWf, g_root, gO.sum(0)
