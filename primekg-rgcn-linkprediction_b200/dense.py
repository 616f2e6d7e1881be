"""Dense relational transform of one RGCN layer (the `h_r @ W_r` products + root term) on tcgen05.

    forward  O  = H @ Wf + X @ root + bias (, ReLU)    H = [H_0 | ... | H_{R-1}]   [N, R*d_in]
    dgrad    gA = (gO * relu') @ [Wf ; root]^T         [N, (R+1)*d_in]
    wgrad    gWf = H^T @ (gO * relu') ,  g_root = X^T @ (gO * relu') ,  g_bias = sum_i (gO * relu')[i]

Replaces the R+1 matmuls per layer of RGCNConv's loop path (reference call sites
src/models/rgcn.py:123, :128) and their autograd transposes (src/train.py:306).  All three run as
hand-written tcgen05/TMEM kernels (csrc/transform.cu) with the bias / ReLU / ReLU-backward / bias-gradient
work fused into the loaders and epilogues.

mode "fp32": operands split into bf16 hi + lo, three products, fp32 accumulate (~1e-5 relative error).
mode "bf16": operands rounded to bf16, fp32 accumulate (the "bf16-transform" mode, tolerance 2e-2).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import ops

MODES = ("fp32", "bf16")


def transform_fwd(H: torch.Tensor, Wf: torch.Tensor, x: torch.Tensor, root: torch.Tensor, bias: torch.Tensor,
                  relu: bool, mode: str) -> torch.Tensor:
    return ops.transform_fwd(H, x, Wf, root, bias, relu, mode)


def transform_dgrad(gO: torch.Tensor, relu_out: Optional[torch.Tensor], Wf: torch.Tensor, root: torch.Tensor,
                    mode: str) -> torch.Tensor:
    """gO [N, d_out] -> [N, (R+1)*d_in] (the last d_in columns are the root / self-loop term)."""
    return ops.transform_dgrad(gO, relu_out, Wf, root, mode)


def transform_wgrad(H: torch.Tensor, x: torch.Tensor, gO: torch.Tensor, relu_out: Optional[torch.Tensor], mode: str
                    ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    return ops.transform_wgrad(H, x, gO, relu_out, mode)
