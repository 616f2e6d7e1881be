"""All-pairs scoring and ranking on device (SURVEY.md §8a `score_all_tails`, §8f rows 1 and 3, BASELINE cfg4).

* ``score_all_pairs``  — DistMult ``(h * r) @ T^T`` (reference src/models/rgcn.py:234-241) or cosine ``(cos + 1) / 2``
  (src/compare_methods.py:384-397; src/medical_validation.py:222-239) over index lists, gathers fused.
* ``rank_true_tails``  — 1-indexed rank of the true tail among all candidates, the quantity the Python loop at
  src/evaluate.py:266-276 extracts with one ``argsort`` per row.

Both are a dense contraction ``A' @ B^T`` over the feature width and run on the tensor cores by default (``method="tc"``):
the prepared query rows become bf16 hi / lo operand planes and the candidate rows the K-major "weight" operand of the
tcgen05 kernel behind ``rgcn_transform_dgrad`` (three bf16 products with fp32 accumulation, 3-7e-6 relative error — the
fp32 mode of the encoder).  ``alpha`` / ``beta`` ride along as one extra K column, so the cosine rescale costs nothing.
The ranking counts come from the score block itself (``rgcn_rank_count``), a block of queries at a time.
``method="simt"`` keeps the fp32 FMA tile kernels of ``csrc/rank.cu`` (every output accumulated in ascending k order with
``fmaf``; they never materialise the [queries, candidates] block when ranking).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib
from .graph import _ptr, _stream

_RANK_BLOCK = 2048       # queries ranked per score block of the tensor-core path (2048 x 30,926 fp32 = 253 MB)


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


def _tc_block(A: torch.Tensor, Bext: torch.Tensor, alpha: float, beta: float) -> torch.Tensor:
    """[rows(A), rows(Bext)] = alpha * A @ Bext[:, :d]^T + beta on the tensor cores.  ``Bext`` is the candidate operand
    from ``_tc_candidates`` ([nb_pad, d + 8] when alpha / beta ride along, else [nb_pad, d])."""
    from . import ops
    d = A.size(1)
    dk = Bext.size(1)
    if dk != d:
        Ae = torch.zeros(A.size(0), dk, dtype=torch.float32, device=A.device)
        torch.mul(A, alpha, out=Ae[:, :d])
        Ae[:, d] = 1.0
        A = Ae
    planes = ops.alloc_planes(A.size(0), dk, "fp32", A.device)
    ops.split_planes(A, planes)
    return ops.transform_dgrad(planes, dk, Bext, None, "fp32")


def _tc_candidates(B: torch.Tensor, b_idx: Optional[torch.Tensor], alpha: float, beta: float) -> torch.Tensor:
    """Candidate rows as the K-major operand: gathered, padded to a multiple of 8 rows with zeros and — when the scores
    are to be rescaled — widened by 8 columns of which the first holds ``beta`` (it meets the constant 1 of the queries)."""
    nb = B.size(0) if b_idx is None else b_idx.numel()
    d = B.size(1)
    ext = alpha != 1.0 or beta != 0.0
    Be = torch.zeros((nb + 7) // 8 * 8, d + (8 if ext else 0), dtype=torch.float32, device=B.device)
    if b_idx is None:
        Be[:nb, :d] = B
    else:
        lib = _lib.load()
        view = Be[:nb, :d]
        _lib.check(lib.rgcn_rows_prepare(_ptr(B), B.stride(0), _ptr(b_idx), nb, d, None, None, 0, _ptr(view), Be.stride(0),
                                         _stream(B.device)), "rgcn_rows_prepare")
    if ext:
        Be[:nb, d] = beta
    return Be


def scores_from_rows(A: torch.Tensor, B: torch.Tensor, b_idx: Optional[torch.Tensor] = None, alpha: float = 1.0,
                     beta: float = 0.0, method: str = "tc") -> torch.Tensor:
    """out[i, j] = alpha * <A[i], B[b_idx[j]]> + beta  (fp32, fused gather of the candidate rows)."""
    lib = _lib.load()
    A = A.detach().to(torch.float32).contiguous()
    B = B.detach().to(torch.float32)
    if B.stride(1) != 1 or B.stride(0) % 4 or B.data_ptr() % 16:
        B = B.contiguous()
    b_idx = None if b_idx is None else b_idx.to(torch.int64).contiguous()
    nb = B.size(0) if b_idx is None else b_idx.numel()
    if method == "tc" and A.size(0) > 0 and nb > 0 and A.size(1) % 4 == 0:
        out = _tc_block(A, _tc_candidates(B, b_idx, alpha, beta), alpha, beta)
        return out[:, :nb]
    if method not in ("tc", "simt"):
        raise ValueError("method must be 'tc' (tensor cores) or 'simt' (fp32 FMA tiles)")
    out = torch.empty(A.size(0), nb, dtype=torch.float32, device=A.device)
    _lib.check(lib.rgcn_allpairs_scores(_ptr(A), A.stride(0), A.size(0), _ptr(B), B.stride(0), _ptr(b_idx), nb,
                                        A.size(1), alpha, beta, _ptr(out), out.stride(0), _stream(A.device)),
               "rgcn_allpairs_scores")
    return out


def score_all_pairs(emb: torch.Tensor, a_idx: torch.Tensor, b_idx: torch.Tensor,
                    rel_vec: Optional[torch.Tensor] = None, cosine: bool = False, method: str = "tc") -> torch.Tensor:
    """[len(a_idx), len(b_idx)] scores.  DistMult: (emb[a] * rel_vec) . emb[b];  cosine: (cos(emb[a], emb[b]) + 1) / 2."""
    if cosine:
        A = _prep(emb, a_idx, None, None, True)
        Bn = _prep(emb, b_idx, None, None, True)
        return scores_from_rows(A, Bn, None, 0.5, 0.5, method=method)
    if rel_vec is not None:
        table = rel_vec.reshape(1, -1)
        A = _prep(emb, a_idx, table, torch.zeros(a_idx.numel(), dtype=torch.int64, device=emb.device), False)
    else:
        A = _prep(emb, a_idx, None, None, False)
    return scores_from_rows(A, emb, b_idx, method=method)


def _query_planes(A: torch.Tensor):
    from . import ops
    planes = ops.alloc_planes(A.size(0), A.size(1), "fp32", A.device)
    ops.split_planes(A, planes)
    return planes


def _candidate_rows(B: torch.Tensor, idx: Optional[torch.Tensor]) -> torch.Tensor:
    """fp32 candidate rows [n, d], gathered through ``idx`` when given (contiguous)."""
    if idx is None:
        return B.contiguous()
    lib = _lib.load()
    out = torch.empty(idx.numel(), B.size(1), dtype=torch.float32, device=B.device)
    _lib.check(lib.rgcn_rows_prepare(_ptr(B), B.stride(0), _ptr(idx), idx.numel(), B.size(1), None, None, 0, _ptr(out),
                                     out.stride(0), _stream(B.device)), "rgcn_rows_prepare")
    return out


def _rank_fused(A: torch.Tensor, B: torch.Tensor, cand: Optional[torch.Tensor], tails: torch.Tensor):
    """Tensor-core sweep whose epilogue counts instead of storing (``rgcn_scores_rank_w``); the threshold comes from the
    diagonal tiles of queries x (their own true tails) with the same K order, so it carries the sweep's own bits."""
    from . import ops
    lib = _lib.load()
    nq, d = A.shape
    C = _candidate_rows(B, cand)
    nb = C.size(0)
    Q = _query_planes(A)
    cand_planes = ops.prepare_weights(C, None, "fp32")
    true_rows = torch.empty(nq, d, dtype=torch.float32, device=A.device)
    _lib.check(lib.rgcn_rows_prepare(_ptr(C), C.stride(0), _ptr(tails), nq, d, None, None, 0, _ptr(true_rows), d,
                                     _stream(A.device)), "rgcn_rows_prepare")
    tail_planes = ops.prepare_weights(true_rows, None, "fp32")
    thr = torch.empty(nq, dtype=torch.float32, device=A.device)
    greater = torch.zeros(nq, dtype=torch.int32, device=A.device)
    equal = torch.zeros(nq, dtype=torch.int32, device=A.device)
    st = _stream(A.device)
    _lib.check(lib.rgcn_scores_diag_w(_ptr(Q[0]), _ptr(Q[1]), Q[0].stride(0), d, _ptr(tail_planes), nq, _ptr(thr), st),
               "rgcn_scores_diag_w")
    _lib.check(lib.rgcn_scores_rank_w(_ptr(Q[0]), _ptr(Q[1]), Q[0].stride(0), d, _ptr(cand_planes), nb, nq, _ptr(thr),
                                      _ptr(tails), _ptr(greater), _ptr(equal), st), "rgcn_scores_rank_w")
    return greater, equal


def topk_from_rows(A: torch.Tensor, B: torch.Tensor, k: int, b_idx: Optional[torch.Tensor] = None, alpha: float = 1.0,
                   beta: float = 0.0) -> Tuple[torch.Tensor, torch.Tensor]:
    """(values [n_a, k], positions [n_a, k]): per row of ``A`` the k best rows of ``B[b_idx]`` by alpha * <a, b> + beta,
    sorted by value (ties: lower position first); positions index ``b_idx`` (or B).  Tensor cores, the score block is
    consumed in the epilogue (``rgcn_scores_topk_w``), k <= 16."""
    from . import ops
    lib = _lib.load()
    if not (1 <= k <= 16):
        raise ValueError("fused top-k serves 1 <= k <= 16")
    A = A.detach().to(torch.float32).contiguous()
    B = B.detach().to(torch.float32)
    if B.stride(1) != 1:
        B = B.contiguous()
    C = _candidate_rows(B, None if b_idx is None else b_idx.to(torch.int64).contiguous())
    na, d = A.shape
    nb = C.size(0)
    if na == 0 or nb == 0 or d % 4:
        raise ValueError("topk_from_rows needs non-empty operands with a feature width that is a multiple of 4")
    Q = _query_planes(A)
    cand_planes = ops.prepare_weights(C, None, "fp32")
    slots = int(lib.rgcn_scores_topk_slots(na, nb))
    dev = A.device
    cv = torch.empty(na, slots, 16, dtype=torch.float32, device=dev)
    ci = torch.empty(na, slots, 16, dtype=torch.int32, device=dev)
    ctr = torch.empty(na, dtype=torch.int32, device=dev)
    ov = torch.empty(na, k, dtype=torch.float32, device=dev)
    oi = torch.empty(na, k, dtype=torch.int64, device=dev)
    _lib.check(lib.rgcn_scores_topk_w(_ptr(Q[0]), _ptr(Q[1]), Q[0].stride(0), d, _ptr(cand_planes), nb, na, int(k),
                                      float(alpha), float(beta), _ptr(cv), _ptr(ci), _ptr(ctr), slots, _ptr(ov), _ptr(oi),
                                      _stream(dev)), "rgcn_scores_topk_w")
    return ov, oi


def topk_all_pairs(emb: torch.Tensor, a_idx: torch.Tensor, b_idx: torch.Tensor, k: int = 10,
                   rel_vec: Optional[torch.Tensor] = None, cosine: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
    """Per row of ``a_idx`` the k best of ``b_idx`` — DistMult ``(emb[a] * rel_vec) . emb[b]`` or cosine ``(cos + 1) / 2`` —
    without the [len(a), len(b)] matrix: (scores [len(a), k], node ids [len(a), k]).  The device form of the per-disease /
    per-drug sweeps + top-k filters of src/compare_methods.py:384-397, src/medical_validation.py:222-239 and
    src/case_studies.py:260-274 (fused L2 normalisation, fused rescale)."""
    b_idx = b_idx.to(torch.int64)
    if cosine:
        A = _prep(emb, a_idx, None, None, True)
        Bn = _prep(emb, b_idx, None, None, True)
        val, pos = topk_from_rows(A, Bn, k, None, 0.5, 0.5)
    else:
        if rel_vec is not None:
            A = _prep(emb, a_idx, rel_vec.reshape(1, -1), torch.zeros(a_idx.numel(), dtype=torch.int64, device=emb.device), False)
        else:
            A = _prep(emb, a_idx, None, None, False)
        val, pos = topk_from_rows(A, emb, k, b_idx)
    ids = torch.where(pos >= 0, b_idx[pos.clamp(min=0)], pos)
    return val, ids


def rank_true_tails(emb: torch.Tensor, rel_table: torch.Tensor, heads: torch.Tensor, rels: torch.Tensor,
                    tails: torch.Tensor, candidates: Optional[torch.Tensor] = None, method: str = "tc"
                    ) -> Tuple[torch.Tensor, torch.Tensor]:
    """(rank, ties): rank[i] = 1 + #{candidates scoring strictly above the true tail} (int64, 1-indexed like
    src/evaluate.py:274); ties[i] = #{other candidates with exactly the true tail's score}.  ``candidates`` = index list
    of admissible tails (default: all entities); ``tails`` then holds POSITIONS in that list.
    ``method``: "tc" = tensor cores, counts taken in the GEMM epilogue (no score block); "tc_block" = tensor cores through
    materialised 2,048-query score blocks (round 1); "simt" = fp32 FMA tiles."""
    lib = _lib.load()
    A = _prep(emb, heads, rel_table, rels, False)
    B = emb.detach().to(torch.float32)
    if B.stride(1) != 1 or B.stride(0) % 4 or B.data_ptr() % 16:
        B = B.contiguous()
    cand = None if candidates is None else candidates.to(torch.int64).contiguous()
    nb = B.size(0) if cand is None else cand.numel()
    nq = A.size(0)
    tails = tails.to(torch.int64).contiguous()
    greater = torch.empty(nq, dtype=torch.int32, device=A.device)
    equal = torch.empty(nq, dtype=torch.int32, device=A.device)
    if method == "tc" and nq > 0 and nb > 0 and A.size(1) % 4 == 0:
        greater, equal = _rank_fused(A, B, cand, tails)
        return greater.to(torch.int64) + 1, equal.to(torch.int64)
    if method in ("tc", "tc_block") and nq > 0 and nb > 0 and A.size(1) % 4 == 0:
        Bext = _tc_candidates(B, cand, 1.0, 0.0)
        # queries per score block: at most _RANK_BLOCK, and at most ~1 GiB of scores for very large candidate sets
        qb = max(128, min(_RANK_BLOCK, (1 << 28) // max(Bext.size(0), 1) // 128 * 128))
        for q0 in range(0, nq, qb):
            q1 = min(q0 + qb, nq)
            S = _tc_block(A[q0:q1], Bext, 1.0, 0.0)
            _lib.check(lib.rgcn_rank_count(_ptr(S), S.stride(0), q1 - q0, nb, _ptr(tails[q0:q1]), None, _ptr(greater[q0:q1]),
                                           _ptr(equal[q0:q1]), _stream(A.device)), "rgcn_rank_count")
        return greater.to(torch.int64) + 1, equal.to(torch.int64)
    if method not in ("tc", "tc_block", "simt"):
        raise ValueError("method must be 'tc' (tensor cores, fused count), 'tc_block' or 'simt' (fp32 FMA tiles)")
    thr = torch.empty(nq, dtype=torch.float32, device=A.device)
    _lib.check(lib.rgcn_allpairs_rank(_ptr(A), A.stride(0), nq, _ptr(B), B.stride(0), _ptr(cand), nb, A.size(1),
                                      _ptr(tails), _ptr(thr), _ptr(greater), _ptr(equal), _stream(A.device)),
               "rgcn_allpairs_rank")
    return greater.to(torch.int64) + 1, equal.to(torch.int64)


def ranking_metrics(ranks: torch.Tensor, k_values=(1, 3, 10, 50, 100), ties: Optional[torch.Tensor] = None,
                    tie_policy: str = "optimistic") -> dict:
    """MRR / mean / median rank / Hits@K as assembled at src/evaluate.py:278-291.

    ``median_rank`` is numpy's median (the MEAN of the two middle order statistics for an even count, as
    ``np.median(ranks)`` at src/evaluate.py:283; ``torch.median`` would return the lower one).
    Ties: ``rank_true_tails`` returns the OPTIMISTIC rank 1 + #{strictly greater}; the reference's position in an
    unstable ``argsort`` is arbitrary inside [rank, rank + ties].  ``tie_policy="mean"`` (needs ``ties``) reports the
    expected rank rank + ties / 2 instead, ``"pessimistic"`` rank + ties."""
    r = ranks.to(torch.float64)
    if tie_policy != "optimistic":
        if ties is None:
            raise ValueError("tie_policy needs the tie counts returned by rank_true_tails")
        if tie_policy == "mean":
            r = r + ties.to(torch.float64) / 2
        elif tie_policy == "pessimistic":
            r = r + ties.to(torch.float64)
        else:
            raise ValueError("tie_policy must be optimistic, mean or pessimistic")
    out = {"mrr": float((1.0 / r).mean()), "mean_rank": float(r.mean()),
           "median_rank": float(torch.quantile(r, 0.5)) if r.numel() else float("nan"), "tie_policy": tie_policy}
    for k in k_values:
        out[f"hits@{k}"] = float((r <= k).double().mean())
    return out
