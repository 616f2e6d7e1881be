"""CPU ORACLE — TEST INFRASTRUCTURE ONLY.  Never imported by the product package.

A plain-PyTorch restatement of the arithmetic on the RGCN hot path of
arnold117/PrimeKG-RGCN-LinkPrediction.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this file,
and only as the checker or the timed CPU baseline — never as a compute fallback.

PARITY PIN STATUS
-----------------
* The graph-convolution operator of the reference is third-party
  ``torch_geometric.nn.RGCNConv`` (imported at reference src/models/rgcn.py:17, built at
  :72-85, called at :123 and :128).  ``torch-geometric`` is pinned only as ``>=2.4.0``
  (reference requirements.txt:2), is NOT vendored under /root/reference and is NOT
  installable here (no wheel, no network).  ``rgcn_conv_ref`` below restates its published
  loop-path algorithm (per relation: mask -> gather -> scatter-mean -> ``@ W_r``; then
  ``+ x @ root + bias``).  For that single operator the parity is therefore
  **"parity unpinned"**: it is cross-checked against an independent dense-adjacency
  formulation (``rgcn_conv_dense_ref``) and ``torch.autograd.gradcheck``, not against PyG.
* Everything around the operator (encoder wiring, DistMult decoder, composite model,
  the train step) IS pinned: ``tests/golden/make_golden.py`` imports the UNMODIFIED
  reference ``src/models/rgcn.py`` with a stub ``torch_geometric.nn`` whose ``RGCNConv``
  is ``RGCNConvRef`` and stores its outputs as fixtures; ``tests/test_oracle.py`` checks
  this file against those fixtures.

Every function cites the reference file:line it follows.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F


# ----------------------------------------------------------------------------------
# Index construction oracle (new work in the product; defined by stable sorts)
# ----------------------------------------------------------------------------------
def csr_oracle(edge_index: torch.Tensor, edge_type: torch.Tensor, num_nodes: int,
               num_relations: int):
    """(dst, relation)-keyed CSR and (src, relation)-keyed transposed CSR.

    The order inside a segment is the ORIGINAL edge order: PyG's
    ``edge_index[:, edge_type == r]`` keeps column order (RGCNConv loop path) and a
    stable sort keeps it inside one key.  Input layout: reference src/preprocess.py:240-261
    (``edge_index[0]`` = source/head, ``edge_index[1]`` = destination/tail, int64).

    Returns int64 tensors (rowptr[N*R+1], col[E], perm[E], rowptr_t[N*R+1], row_t[E], perm_t[E]).
    """
    src, dst = edge_index[0].long(), edge_index[1].long()
    rel = edge_type.long()
    E = src.numel()
    if E:
        if int(src.min()) < 0 or int(src.max()) >= num_nodes or int(dst.min()) < 0 or int(dst.max()) >= num_nodes:
            raise IndexError("edge_index out of range")
        if int(rel.min()) < 0 or int(rel.max()) >= num_relations:
            raise IndexError("edge_type out of range")
    nk = num_nodes * num_relations

    def one(major, minor):
        key = major * num_relations + rel
        perm = torch.sort(key, stable=True).indices
        cnt = torch.bincount(key, minlength=nk)
        rowptr = torch.zeros(nk + 1, dtype=torch.int64)
        rowptr[1:] = torch.cumsum(cnt, 0)
        return rowptr, minor[perm], perm

    rowptr, col, perm = one(dst, src)
    rowptr_t, row_t, perm_t = one(src, dst)
    return rowptr, col, perm, rowptr_t, row_t, perm_t


# ----------------------------------------------------------------------------------
# RGCNConv (PyG, not in tree) — loop path
# ----------------------------------------------------------------------------------
def glorot_(t: Optional[torch.Tensor]) -> None:
    """torch_geometric.nn.inits.glorot: U(-a, a), a = sqrt(6 / (size(-2) + size(-1)))."""
    if t is not None:
        a = math.sqrt(6.0 / (t.size(-2) + t.size(-1)))
        t.data.uniform_(-a, a)


def rgcn_conv_ref(x, edge_index, edge_type, weight, root, bias, comp=None):
    """RGCNConv.forward(x, edge_index, edge_type), loop path (call sites: reference
    src/models/rgcn.py:123, :128).  x'_i = root^T x_i + b + sum_r mean_{j in N_r(i)} W_r^T x_j,
    the mean taken per (destination, relation) with an empty neighbourhood giving 0."""
    N = x.size(0)
    if comp is not None:                       # basis decomposition, num_bases = weight.size(0)
        R = comp.size(0)
        W = (comp @ weight.view(weight.size(0), -1)).view(R, weight.size(1), weight.size(2))
    else:
        R = weight.size(0)
        W = weight
    out = torch.zeros(N, W.size(2), dtype=x.dtype, device=x.device)
    for r in range(R):
        m = edge_type == r
        src, dst = edge_index[0][m], edge_index[1][m]      # keeps original edge order
        xj = x.index_select(0, src)
        s = torch.zeros(N, x.size(1), dtype=x.dtype, device=x.device).index_add_(0, dst, xj)
        cnt = torch.zeros(N, dtype=x.dtype, device=x.device).index_add_(
            0, dst, torch.ones(dst.numel(), dtype=x.dtype, device=x.device)).clamp_(min=1)
        h = s / cnt.unsqueeze(1)
        out = out + h @ W[r]                                # aggregate first, then transform
    out = out + x @ root
    out = out + bias
    return out


def rgcn_conv_dense_ref(x, edge_index, edge_type, weight, root, bias, comp=None):
    """Independent formulation for cross-checking the restatement on tiny graphs:
    dense multiplicity matrices A_r[i, j] = #edges j->i of type r, row-normalised."""
    N = x.size(0)
    if comp is not None:
        R = comp.size(0)
        W = torch.einsum("rb,bio->rio", comp, weight)
    else:
        R = weight.size(0)
        W = weight
    out = x @ root + bias
    for r in range(R):
        A = torch.zeros(N, N, dtype=x.dtype)
        m = edge_type == r
        A.index_put_((edge_index[1][m], edge_index[0][m]),
                     torch.ones(int(m.sum()), dtype=x.dtype), accumulate=True)
        A = A / A.sum(1, keepdim=True).clamp(min=1)
        out = out + A @ (x @ W[r])                          # transform-first order on purpose
    return out


class RGCNConvRef(nn.Module):
    """Parameter layout and init order of PyG's RGCNConv (weight, comp, root, bias)."""

    def __init__(self, in_channels, out_channels, num_relations, num_bases=None):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.num_relations, self.num_bases = num_relations, num_bases
        if num_bases is not None:
            self.weight = nn.Parameter(torch.empty(num_bases, in_channels, out_channels))
            self.comp = nn.Parameter(torch.empty(num_relations, num_bases))
        else:
            self.weight = nn.Parameter(torch.empty(num_relations, in_channels, out_channels))
            self.register_parameter("comp", None)
        self.root = nn.Parameter(torch.empty(in_channels, out_channels))
        self.bias = nn.Parameter(torch.empty(out_channels))
        self.reset_parameters()

    def reset_parameters(self):
        glorot_(self.weight)
        glorot_(self.comp)
        glorot_(self.root)
        self.bias.data.zero_()

    def forward(self, x, edge_index, edge_type):
        assert edge_type is not None
        return rgcn_conv_ref(x, edge_index, edge_type, self.weight, self.root, self.bias, self.comp)


# ----------------------------------------------------------------------------------
# Model file restatement (reference src/models/rgcn.py)
# ----------------------------------------------------------------------------------
class EncoderRef(nn.Module):
    """DrugDiseaseRGCN — reference src/models/rgcn.py:51-130."""

    def __init__(self, num_nodes, num_relations, embedding_dim=64, hidden_dim=128, dropout=0.5,
                 num_bases=None, num_layers=2):
        super().__init__()
        self.node_embeddings = nn.Embedding(num_nodes, embedding_dim)            # :68
        self.conv1 = RGCNConvRef(embedding_dim, hidden_dim, num_relations, num_bases)   # :72-77
        self.conv2 = RGCNConvRef(hidden_dim, hidden_dim, num_relations, num_bases)      # :80-85
        # extension pattern of reference guide/MODEL_ARCHITECTURE.md:245-249 (cfg5 uses 3 layers)
        self.extra = nn.ModuleList(
            [RGCNConvRef(hidden_dim, hidden_dim, num_relations, num_bases) for _ in range(num_layers - 2)])
        self.dropout = nn.Dropout(dropout)                                        # :88
        nn.init.xavier_uniform_(self.node_embeddings.weight)                      # :93-95

    def forward(self, edge_index, edge_type, node_indices=None):
        x = self.node_embeddings.weight if node_indices is None else self.node_embeddings(node_indices)  # :117-120
        x = self.conv1(x, edge_index, edge_type)          # :123
        x = F.relu(x)                                     # :124
        x = self.dropout(x)                               # :125
        x = self.conv2(x, edge_index, edge_type)          # :128
        for conv in self.extra:
            x = self.dropout(F.relu(x))
            x = conv(x, edge_index, edge_type)
        return x


class DecoderRef(nn.Module):
    """LinkPredictor (DistMult) — reference src/models/rgcn.py:165-243."""

    def __init__(self, num_relations, embedding_dim, dropout=0.0):
        super().__init__()
        self.relation_embeddings = nn.Embedding(num_relations, embedding_dim)     # :177
        self.dropout = nn.Dropout(dropout)                                        # :180
        nn.init.xavier_uniform_(self.relation_embeddings.weight)                  # :185-187

    def forward(self, head_embeddings, tail_embeddings, relation_types):
        r = self.dropout(self.relation_embeddings(relation_types))                # :207-208
        return torch.sum(head_embeddings * r * tail_embeddings, dim=1)            # :211

    def score_all_tails(self, head_embeddings, relation_types, all_tail_embeddings):
        hr = head_embeddings * self.relation_embeddings(relation_types)           # :235-238
        return hr @ all_tail_embeddings.t()                                       # :241


class ModelRef(nn.Module):
    """DrugDiseaseModel — reference src/models/rgcn.py:267-331."""

    def __init__(self, num_nodes, num_relations, embedding_dim=64, hidden_dim=128, dropout=0.5,
                 decoder_dropout=0.0, num_bases=None, num_layers=2):
        super().__init__()
        self.encoder = EncoderRef(num_nodes, num_relations, embedding_dim, hidden_dim, dropout,
                                  num_bases, num_layers)
        self.decoder = DecoderRef(num_relations, hidden_dim, decoder_dropout)

    def forward(self, edge_index, edge_type, head_indices, tail_indices, relation_types):
        h = self.encoder(edge_index, edge_type)                                   # :322
        return self.decoder(h[head_indices], h[tail_indices], relation_types)     # :325-329


def train_step_ref(model, edge_index, edge_type, heads, tails, rels, labels):
    """One hot-path step: reference src/train.py:291-306 (forward, BCEWithLogits, backward).
    Returns (loss, scores); grads are left in ``.grad`` of the parameters."""
    scores = model(edge_index, edge_type, heads, tails, rels)
    loss = F.binary_cross_entropy_with_logits(scores, labels)                     # train.py:139, :300
    loss.backward()                                                               # train.py:306
    return loss.detach(), scores.detach()


def negative_batch_ref(pos_head, pos_tail, pos_rel, num_nodes: int, num_neg: int = 1, generator=None):
    """The mini-batch the reference's loop assembles: ``NegativeSampler.sample`` (src/train.py:59-97: repeat_interleave,
    corrupt the head where rand < 0.5 else the tail, replacement uniform over all nodes) followed by the concatenation
    and labels of src/train.py:281-288.  Returns (heads, tails, rels, labels).  The device sampler draws from another
    random stream, so parity with this restatement is distributional (layout, rates, ranges), not element-wise."""
    n = pos_head.numel()
    nh, nt, nr = (t.repeat_interleave(num_neg) for t in (pos_head, pos_tail, pos_rel))
    total = n * num_neg
    corrupt_head = torch.rand(total, generator=generator) < 0.5
    ent = torch.randint(0, num_nodes, (total,), generator=generator)
    nh = torch.where(corrupt_head, ent, nh)
    nt = torch.where(~corrupt_head, ent, nt)
    labels = torch.cat([torch.ones(n), torch.zeros(total)])
    return torch.cat([pos_head, nh]), torch.cat([pos_tail, nt]), torch.cat([pos_rel, nr]), labels


def accuracy_count_ref(scores: torch.Tensor, labels: torch.Tensor) -> int:
    """Number of correct predictions as counted at src/train.py:321-322."""
    return int(((torch.sigmoid(scores) > 0.5).float() == labels).sum().item())


# ----------------------------------------------------------------------------------
# Ranking / all-pairs scoring (cfg4 and the evaluate.py inner loop)
# ----------------------------------------------------------------------------------
def rank_of_true_tail_ref(scores: torch.Tensor, true_tail: torch.Tensor):
    """1-indexed rank of the true tail among all candidates — reference src/evaluate.py:266-276
    (descending argsort, position of the true tail).  ``torch.argsort`` is unstable so exact ties
    are arbitrary in the reference; this oracle returns the two deterministic bounds
    (optimistic = 1 + #greater, pessimistic = #greater-or-equal) which bracket it."""
    s_true = scores.gather(1, true_tail.view(-1, 1))
    greater = (scores > s_true).sum(1)
    geq = (scores >= s_true).sum(1)
    return greater + 1, geq


def distmult_allpairs_ref(emb, head_idx, tail_idx, rel_vec):
    """(h ⊙ r) @ T^T over a head set × tail set — reference src/models/rgcn.py:234-241."""
    return (emb[head_idx] * rel_vec) @ emb[tail_idx].t()


def cosine_allpairs_ref(emb, a_idx, b_idx):
    """(cos + 1) / 2 on L2-normalised rows — reference src/compare_methods.py:384-397."""
    a = emb[a_idx]
    b = emb[b_idx]
    a = a / a.norm(dim=1, keepdim=True)
    b = b / b.norm(dim=1, keepdim=True)
    return (a @ b.t() + 1) / 2
