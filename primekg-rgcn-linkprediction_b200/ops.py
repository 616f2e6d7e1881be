"""Thin host wrappers: one Python function per C-ABI entry point of ``include/rgcn_b200.h``.

Every function launches hand-written sm_100a kernels on torch's current stream and returns
torch tensors; nothing here computes on the CPU or through another library.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from .graph import RelGraph, _ptr, _stream


def _f32c(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: the RGCN B200 path has no CPU implementation")
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32, got {t.dtype}")
    if t.dim() != 2:
        raise ValueError(f"{name} must be 2-D")
    if t.stride(1) != 1 or t.stride(0) % 4 != 0 or t.data_ptr() % 16 != 0:
        t = t.contiguous()
    return t


def aggregate_fwd(g: RelGraph, x: torch.Tensor, out_bf16: bool = False,
                  comp: Optional[torch.Tensor] = None) -> torch.Tensor:
    """H[i, r*d:(r+1)*d] = mean_{j in N_r(i)} x[j]  (or the basis-mixed Z when ``comp`` [R, B] is given)."""
    lib = _lib.load()
    x = _f32c(x, "x")
    if x.size(0) != g.n_src:
        raise ValueError(f"x has {x.size(0)} rows, the graph gathers from {g.n_src}")
    d = x.size(1)
    if d % 4 != 0 or d > 1024:
        raise ValueError("feature width must be a multiple of 4 and at most 1024")
    blocks = g.R if comp is None else int(comp.size(1))
    if comp is not None:
        comp = comp.detach().to(torch.float32).contiguous()
        if comp.size(0) != g.R:
            raise ValueError("comp must have one row per relation")
    H = torch.empty(g.n_dst, blocks * d, dtype=torch.bfloat16 if out_bf16 else torch.float32, device=x.device)
    ws = g.fwd.workspace(d)
    _lib.check(lib.rgcn_aggregate_fwd(g.fwd.ref, _ptr(x), x.stride(0), d, _ptr(comp), 0 if comp is None else blocks,
                                      _ptr(H), H.stride(0), int(out_bf16), _ptr(ws),
                                      0 if ws is None else ws.numel() * 4, _stream(x.device)), "rgcn_aggregate_fwd")
    return H


def aggregate_bwd(g: RelGraph, gH: torch.Tensor, d: int, init: Optional[torch.Tensor] = None) -> torch.Tensor:
    """gX[j] = init[j] + sum_r sum_{(j->i, r)} gH[i, r*d:(r+1)*d] / max(|N_r(i)|, 1)  over the transposed CSR."""
    lib = _lib.load()
    gH = _f32c(gH, "gH")
    if gH.size(0) != g.n_dst or gH.size(1) < g.R * d:
        raise ValueError("gH must be [n_dst, >= R*d]")
    if init is not None:
        init = _f32c(init, "init")
        if init.size(0) != g.n_src or init.size(1) < d:
            raise ValueError("init must be [n_src, >= d]")
    gX = torch.empty(g.n_src, d, dtype=torch.float32, device=gH.device)
    ws = g.bwd.workspace(d)
    _lib.check(lib.rgcn_aggregate_bwd(g.bwd.ref, _ptr(gH), gH.stride(0), d, _ptr(init),
                                      0 if init is None else init.stride(0), _ptr(gX), gX.stride(0), _ptr(ws),
                                      0 if ws is None else ws.numel() * 4, _stream(gH.device)), "rgcn_aggregate_bwd")
    return gX


def _idx(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    return t.to(torch.int64).contiguous()


def check_pairs(head, tail, rel, n_nodes: int, n_rel: int) -> None:
    """Raise IndexError on an out-of-range pair (synchronises; used outside the training hot loop)."""
    lib = _lib.load()
    flag = torch.zeros(1, dtype=torch.int32, device=head.device)
    _lib.check(lib.rgcn_check_pairs(_ptr(head), _ptr(tail), _ptr(rel), head.numel(), n_nodes, n_rel, _ptr(flag),
                                    _stream(head.device)), "rgcn_check_pairs")
    if int(flag.item()):
        raise IndexError("head/tail/relation index out of range")


def _rows(t, name):
    return None if t is None else _f32c(t, name).contiguous()


def distmult_fwd(emb_h: torch.Tensor, emb_t: torch.Tensor, head, tail, rel, rel_table: Optional[torch.Tensor],
                 rel_rows: Optional[torch.Tensor]) -> torch.Tensor:
    """score[p] = sum_k emb_h[hp, k] * r_p[k] * emb_t[tp, k]; hp = head[p] (or p when head is None)."""
    lib = _lib.load()
    emb_h, emb_t = _f32c(emb_h, "head embeddings"), _f32c(emb_t, "tail embeddings")
    head = None if head is None else _idx(head, "head")
    tail = None if tail is None else _idx(tail, "tail")
    rel = None if rel is None else _idx(rel, "rel")
    n = head.numel() if head is not None else emb_h.size(0)
    d = emb_h.size(1)
    if emb_t.size(1) != d:
        raise ValueError("head and tail embeddings differ in width")
    rel_rows, rel_table = _rows(rel_rows, "rel_rows"), _rows(rel_table, "rel_table")
    score = torch.empty(n, dtype=torch.float32, device=emb_h.device)
    _lib.check(lib.rgcn_distmult_fwd(_ptr(emb_h), emb_h.stride(0), _ptr(emb_t), emb_t.stride(0), _ptr(head),
                                     _ptr(tail), _ptr(rel), _ptr(rel_table), _ptr(rel_rows), n, d, _ptr(score),
                                     _stream(emb_h.device)), "rgcn_distmult_fwd")
    return score


def distmult_bwd(emb_h, emb_t, head, tail, rel, rel_table, rel_rows, g_score, need_rel_table_grad: bool):
    """Returns (g_h, g_t, g_rel_table | None, g_rel_rows | None).  With index arrays and emb_h is emb_t the
    two row gradients are accumulated into ONE dense [N, d] buffer (returned as g_h, g_t = None)."""
    lib = _lib.load()
    shared = emb_h is emb_t and head is not None and tail is not None
    emb_h = _f32c(emb_h, "head embeddings")
    emb_t = emb_h if shared else _f32c(emb_t, "tail embeddings")
    head = None if head is None else _idx(head, "head")
    tail = None if tail is None else _idx(tail, "tail")
    rel = None if rel is None else _idx(rel, "rel")
    g_score = g_score.to(torch.float32).contiguous()
    n = g_score.numel()
    d = emb_h.size(1)
    dev = emb_h.device
    mk = torch.zeros if head is not None else torch.empty
    g_h = mk(emb_h.size(0), d, dtype=torch.float32, device=dev)
    if shared:
        g_t = g_h
    else:
        mk = torch.zeros if tail is not None else torch.empty
        g_t = mk(emb_t.size(0), d, dtype=torch.float32, device=dev)
    rel_rows, rel_table = _rows(rel_rows, "rel_rows"), _rows(rel_table, "rel_table")
    g_rows = torch.empty(n, d, dtype=torch.float32, device=dev) if rel_rows is not None else None
    g_tab = torch.zeros_like(rel_table) if (rel_rows is None and need_rel_table_grad) else None
    _lib.check(lib.rgcn_distmult_bwd(_ptr(emb_h), emb_h.stride(0), _ptr(emb_t), emb_t.stride(0), _ptr(head),
                                     _ptr(tail), _ptr(rel), _ptr(rel_table), _ptr(rel_rows), _ptr(g_score), n, d,
                                     _ptr(g_h), g_h.stride(0), _ptr(g_t), g_t.stride(0), _ptr(g_tab), _ptr(g_rows),
                                     _stream(dev)), "rgcn_distmult_bwd")
    return g_h, (None if shared else g_t), g_tab, g_rows
