"""Row-sparse output gradients: the hand-over between the decoder's backward and the last encoder layer's backward.

The training step of the reference (src/train.py:291-306) reads only ``node_embeddings[head]`` / ``[tail]`` of the
encoder output (src/models/rgcn.py:325-326), so the gradient autograd hands to the last ``RGCNConv`` is an ``[N, d]``
matrix that is zero outside <= 2 * batch rows.  ``_DistMultGather.backward`` still returns that dense matrix (autograd's
contract; the engine may add other contributions to it), and *announces* here which rows it wrote.  The layer's backward
takes the row-sparse path (csrc/rowsparse.cu) only if the tensor it receives is that very buffer, unmodified: same
storage, same shape, same in-place version.  Anything else — another consumer of the embeddings whose gradient the
engine added, a hook that rewrote it — fails the check and the dense path runs.  ``PRIMEKG_RGCN_SPARSE_BWD=0`` turns
the hand-over off.
"""
from __future__ import annotations

import os
from typing import Optional

import torch

MAX_FRACTION = 0.5     # longer row lists (relative to the node count) take the dense backward
stats = {"claimed": 0, "declined": 0}      # how often the last layer took the compact / the dense backward
_announced = None      # (dense gradient tensor, its _version when announced, int64 device list of the rows written,
                       #  int32 node -> first-position map of that list | None)


def enabled() -> bool:
    return os.environ.get("PRIMEKG_RGCN_SPARSE_BWD", "1") != "0"


def planes_enabled() -> bool:
    """Second hand-over (the walk writes the upstream layer's masked planes).  OPT-IN, PRIMEKG_RGCN_PLANES_HANDOVER=1:
    measured on the B200 (cfg2) the walk with the second output takes 68 us against 38 us + a 14-21 us conversion pass
    — the extra registers cost the latency-bound walk a resident block — so the step is 0.427 ms with it, 0.402 without."""
    return enabled() and os.environ.get("PRIMEKG_RGCN_PLANES_HANDOVER", "0") == "1"


def announce(dense: torch.Tensor, rows: torch.Tensor, slot: Optional[torch.Tensor] = None) -> None:
    """``dense`` is zero outside ``rows`` (duplicates allowed).  Holding the tensor keeps its storage from being reused
    while the announcement stands, which is what makes the pointer comparison in ``claim`` sound.  ``slot``: the
    node -> first-position map of ``rows`` when the announcer has built it (``rgcn_link_loss_bwd_rows``)."""
    global _announced
    _announced = (dense, dense._version, rows, slot) if enabled() else None


def claim(grad: torch.Tensor):
    """(rows, slot | None) if ``grad`` is the announced buffer, untouched, and short enough to pay off; else None.
    An announcement is consumed by the first claim attempt."""
    global _announced
    a, _announced = _announced, None
    if a is None:
        return None
    dense, version, rows, slot = a
    same = (grad.data_ptr() == dense.data_ptr() and grad.shape == dense.shape and grad.stride() == dense.stride()
            and grad.dtype == dense.dtype and grad._version == version and dense._version == version)
    if not same or rows.numel() == 0 or rows.numel() > MAX_FRACTION * grad.size(0):
        stats["declined"] += 1
        return None
    stats["claimed"] += 1
    return rows, slot


# ---- second hand-over: a layer's backward walk has already written its input gradient, masked for the layer upstream,
#      as that layer's operand planes (rgcn_masked_planes_out); the upstream backward then skips its conversion pass ----
_planes = None          # (g_x tensor, version, mask data_ptr, scale, mode, (planes, colsum))
plane_stats = {"claimed": 0, "declined": 0}


def announce_planes(g_x: torch.Tensor, mask: torch.Tensor, scale: float, mode: str, payload) -> None:
    global _planes
    _planes = (g_x, g_x._version, mask.data_ptr(), tuple(mask.shape), float(scale), mode, payload) if planes_enabled() else None


def claim_planes(grad: torch.Tensor, mask: Optional[torch.Tensor], scale: float, mode: str):
    """((hi, lo | None), colsum) if ``grad`` is the announced buffer, untouched, and was masked with this very ``mask``
    tensor and scale in this mode; else None.  Consumed by the first attempt."""
    global _planes
    a, _planes = _planes, None
    if a is None or mask is None:
        return None
    g_x, version, mptr, mshape, ascale, amode, payload = a
    ok = (grad.data_ptr() == g_x.data_ptr() and grad.shape == g_x.shape and grad.stride() == g_x.stride()
          and grad._version == version and g_x._version == version and mask.data_ptr() == mptr
          and tuple(mask.shape) == mshape and float(scale) == ascale and mode == amode)
    plane_stats["claimed" if ok else "declined"] += 1
    return payload if ok else None


# ---- third hand-over (forward): a layer's transform has also left its output rounded to bf16; the next layer's walk
#      gathers that copy (bf16-transform mode) ----
_bf16 = None            # (output tensor, version, bf16 copy)


def announce_bf16(out: torch.Tensor, out16: torch.Tensor) -> None:
    global _bf16
    _bf16 = (out, out._version, out16)


def claim_bf16(x: torch.Tensor) -> Optional[torch.Tensor]:
    """The announced bf16 copy if ``x`` is the announced tensor, untouched; consumed by the first attempt."""
    global _bf16
    a, _bf16 = _bf16, None
    if a is None:
        return None
    out, version, out16 = a
    ok = (x.data_ptr() == out.data_ptr() and x.shape == out.shape and x.stride() == out.stride() and x._version == version
          and out._version == version)
    return out16 if ok else None


# ---- fourth hand-over (forward -> decoder backward): the encoder output was computed at a row list only (listed-rows
#      last layer); its consumer's backward then need not define the gradient outside those rows ----
_listed_out = None      # (data_ptr, shape, version)


def mark_listed_output(out: torch.Tensor) -> None:
    global _listed_out
    _listed_out = (out.data_ptr(), tuple(out.shape), out._version)


def unmark_listed_output() -> None:
    """Called by every layer forward that computes ALL rows: a stale mark must never match a recycled allocation."""
    global _listed_out
    _listed_out = None


def is_listed_output(emb: torch.Tensor) -> bool:
    """True if ``emb`` is the marked tensor, untouched; the mark is consumed by the first attempt."""
    global _listed_out
    a, _listed_out = _listed_out, None
    return a is not None and emb.data_ptr() == a[0] and tuple(emb.shape) == a[1] and emb._version == a[2]


def clear() -> None:
    global _announced, _planes, _bf16, _listed_out
    _announced = None
    _planes = None
    _bf16 = None
    _listed_out = None
