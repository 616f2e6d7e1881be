"""All-pairs scoring and ranking on device (SURVEY.md §8a `score_all_tails`, §8f rows 1 and 3, BASELINE cfg4).

* ``score_all_pairs``  — DistMult ``(h * r) @ T^T`` (reference src/models/rgcn.py:234-241) or cosine ``(cos + 1) / 2``
  (src/compare_methods.py:384-397; src/medical_validation.py:222-239) over index lists, gathers fused.
* ``rank_true_tails``  — 1-indexed rank of the true tail among all candidates, the quantity the Python loop at
  src/evaluate.py:266-276 extracts with one ``argsort`` per row, computed without the [batch, N] score matrix.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib
from .graph import _ptr, _stream


def _prep(emb: torch.Tensor, idx: Optional[torch.Tensor], rel_table: Optional[torch.Tensor],
          rel: Optional[torch.Tensor], normalize: bool) -> torch.Tensor:
    lib = _lib.load()
    if not emb.is_cuda:
        raise RuntimeError("all-pairs scoring needs CUDA tensors: there is no CPU implementation of this path")
    emb = emb.detach().to(torch.float32)
    if emb.stride(1) != 1:
        emb = emb.contiguous()
    n = emb.size(0) if idx is None else idx.numel()
    d = emb.size(1)
    idx = None if idx is None else idx.to(torch.int64).contiguous()
    if rel is not None:
        rel = rel.to(torch.int64).contiguous()
        rel_table = rel_table.detach().to(torch.float32).contiguous()
    out = torch.empty(n, d, dtype=torch.float32, device=emb.device)
    _lib.check(lib.rgcn_rows_prepare(_ptr(emb), emb.stride(0), _ptr(idx), n, d, _ptr(rel_table) if rel is not None else None,
                                     _ptr(rel), int(normalize), _ptr(out), out.stride(0), _stream(emb.device)),
               "rgcn_rows_prepare")
    return out


def scores_from_rows(A: torch.Tensor, B: torch.Tensor, b_idx: Optional[torch.Tensor] = None, alpha: float = 1.0,
                     beta: float = 0.0) -> torch.Tensor:
    """out[i, j] = alpha * <A[i], B[b_idx[j]]> + beta  (fp32, fused gather of the candidate rows)."""
    lib = _lib.load()
    A = A.detach().to(torch.float32).contiguous()
    B = B.detach().to(torch.float32)
    if B.stride(1) != 1 or B.stride(0) % 4 or B.data_ptr() % 16:
        B = B.contiguous()
    b_idx = None if b_idx is None else b_idx.to(torch.int64).contiguous()
    nb = B.size(0) if b_idx is None else b_idx.numel()
    out = torch.empty(A.size(0), nb, dtype=torch.float32, device=A.device)
    _lib.check(lib.rgcn_allpairs_scores(_ptr(A), A.stride(0), A.size(0), _ptr(B), B.stride(0), _ptr(b_idx), nb,
                                        A.size(1), alpha, beta, _ptr(out), out.stride(0), _stream(A.device)),
               "rgcn_allpairs_scores")
    return out


def score_all_pairs(emb: torch.Tensor, a_idx: torch.Tensor, b_idx: torch.Tensor,
                    rel_vec: Optional[torch.Tensor] = None, cosine: bool = False) -> torch.Tensor:
    """[len(a_idx), len(b_idx)] scores.  DistMult: (emb[a] * rel_vec) . emb[b];  cosine: (cos(emb[a], emb[b]) + 1) / 2."""
    if cosine:
        A = _prep(emb, a_idx, None, None, True)
        Bn = _prep(emb, b_idx, None, None, True)
        return scores_from_rows(A, Bn, None, 0.5, 0.5)
    if rel_vec is not None:
        table = rel_vec.reshape(1, -1)
        A = _prep(emb, a_idx, table, torch.zeros(a_idx.numel(), dtype=torch.int64, device=emb.device), False)
    else:
        A = _prep(emb, a_idx, None, None, False)
    return scores_from_rows(A, emb, b_idx)


def rank_true_tails(emb: torch.Tensor, rel_table: torch.Tensor, heads: torch.Tensor, rels: torch.Tensor,
                    tails: torch.Tensor, candidates: Optional[torch.Tensor] = None
                    ) -> Tuple[torch.Tensor, torch.Tensor]:
    """(rank, ties): rank[i] = 1 + #{candidates scoring strictly above the true tail} (int64, 1-indexed like
    src/evaluate.py:274); ties[i] = #{other candidates with exactly the true tail's score}.  ``candidates`` = index list
    of admissible tails (default: all entities); ``tails`` then holds POSITIONS in that list."""
    lib = _lib.load()
    A = _prep(emb, heads, rel_table, rels, False)
    B = emb.detach().to(torch.float32)
    if B.stride(1) != 1 or B.stride(0) % 4 or B.data_ptr() % 16:
        B = B.contiguous()
    cand = None if candidates is None else candidates.to(torch.int64).contiguous()
    nb = B.size(0) if cand is None else cand.numel()
    nq = A.size(0)
    tails = tails.to(torch.int64).contiguous()
    thr = torch.empty(nq, dtype=torch.float32, device=A.device)
    greater = torch.empty(nq, dtype=torch.int32, device=A.device)
    equal = torch.empty(nq, dtype=torch.int32, device=A.device)
    _lib.check(lib.rgcn_allpairs_rank(_ptr(A), A.stride(0), nq, _ptr(B), B.stride(0), _ptr(cand), nb, A.size(1),
                                      _ptr(tails), _ptr(thr), _ptr(greater), _ptr(equal), _stream(A.device)),
               "rgcn_allpairs_rank")
    return greater.to(torch.int64) + 1, equal.to(torch.int64)


def ranking_metrics(ranks: torch.Tensor, k_values=(1, 3, 10, 50, 100)) -> dict:
    """MRR / mean / median rank / Hits@K as assembled at src/evaluate.py:278-291."""
    r = ranks.to(torch.float64)
    out = {"mrr": float((1.0 / r).mean()), "mean_rank": float(r.mean()), "median_rank": float(r.median())}
    for k in k_values:
        out[f"hits@{k}"] = float((ranks <= k).double().mean())
    return out
