"""Synthetic PrimeKG-shaped knowledge graphs.

The reference's real graph tensors (``train_data.pt`` / ``full_graph.pt``) are not shipped
(reference .MISSING_LARGE_BLOBS), so every measured workload is generated here, in the
on-disk format of reference src/preprocess.py:228-261: ``edge_index [2, E] int64`` with
row 0 = source/head and row 1 = destination/tail, every undirected edge emitted as two
CONSECUTIVE columns (a->b), (b->a) of the same type, multi-edges kept; ``edge_type [E] int64``.

Endpoints are drawn inside node-type blocks with a heavy tail (idx = lo + floor(n * u**2.5))
so hub nodes exist, as in the real power-law graph.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Tuple

import torch

# node-type blocks of the 3-relation subgraph: reference README.md:45-48, data/processed/mappings.pt
CFG1_BLOCKS = {"disease": (0, 5593), "drug": (5593, 11875), "gene": (11875, 30926)}
# relation -> (block of endpoint a, block of endpoint b, share of undirected edges);
# shares follow reference data/processed/statistics.csv:2 (51,306 / 160,822 / 642,150 rows)
CFG1_RELS = [("drug", "gene", 51306), ("gene", "disease", 160822), ("gene", "gene", 642150)]

# full PrimeKG: 10 node types, 129,375 nodes, 30 relations, 8,100,498 directed edges
# (reference data/processed/statistics.csv:2 total_edges / total_node_types / total_relation_types)
CFG3_BLOCK_SIZES = [28642, 27671, 17080, 15311, 14035, 11169, 7957, 4176, 2516, 818]


@dataclass
class KG:
    edge_index: torch.Tensor      # [2, E] int64
    edge_type: torch.Tensor       # [E] int64
    num_nodes: int
    num_relations: int
    blocks: Dict[str, Tuple[int, int]]

    @property
    def num_edges(self) -> int:
        return int(self.edge_type.numel())

    def as_dict(self):
        """The dict layout of reference src/preprocess.py:256-261."""
        return {"edge_index": self.edge_index, "edge_type": self.edge_type,
                "num_nodes": self.num_nodes, "num_relations": self.num_relations}


def _draw(n_lo: int, n_hi: int, count: int, g: torch.Generator, power: float) -> torch.Tensor:
    u = torch.rand(count, generator=g, dtype=torch.float64)
    n = n_hi - n_lo
    return (n_lo + torch.floor(n * u.pow(power)).clamp_(max=n - 1)).to(torch.int64)


def _emit(pairs: List[Tuple[torch.Tensor, torch.Tensor, int]], num_nodes, num_relations, blocks) -> KG:
    src, dst, typ = [], [], []
    for a, b, r in pairs:
        # consecutive (a->b),(b->a) columns — reference src/preprocess.py:228-234
        src.append(torch.stack([a, b], 1).reshape(-1))
        dst.append(torch.stack([b, a], 1).reshape(-1))
        typ.append(torch.full((2 * a.numel(),), r, dtype=torch.int64))
    ei = torch.stack([torch.cat(src), torch.cat(dst)], 0).contiguous()
    return KG(ei, torch.cat(typ).contiguous(), num_nodes, num_relations, blocks)


def primekg_subgraph(num_directed_edges: int = 849_456, seed: int = 42, power: float = 2.5) -> KG:
    """cfg1 / cfg2 graph: 30,926 nodes, 3 relations, E directed edges (default 849,456)."""
    g = torch.Generator().manual_seed(seed)
    und = num_directed_edges // 2
    tot = sum(c for _, _, c in CFG1_RELS)
    counts = [und * c // tot for _, _, c in CFG1_RELS]
    counts[-1] += und - sum(counts)
    pairs = []
    for r, ((ba, bb, _), c) in enumerate(zip(CFG1_RELS, counts)):
        a = _draw(*CFG1_BLOCKS[ba], c, g, power)
        b = _draw(*CFG1_BLOCKS[bb], c, g, power)
        pairs.append((a, b, r))
    return _emit(pairs, 30_926, 3, dict(CFG1_BLOCKS))


def primekg_full(num_directed_edges: int = 8_100_498, seed: int = 42, power: float = 2.5) -> KG:
    """cfg3 graph: 129,375 nodes in 10 type blocks, 30 relations, ~8.1 M directed edges."""
    g = torch.Generator().manual_seed(seed)
    bounds, lo = [], 0
    for s in CFG3_BLOCK_SIZES:
        bounds.append((lo, lo + s))
        lo += s
    R = 30
    und = num_directed_edges // 2
    w = torch.tensor([1.0 / (r + 1) for r in range(R)], dtype=torch.float64)
    counts = torch.floor(w / w.sum() * und).to(torch.int64).tolist()
    counts[0] += und - sum(counts)
    pairs = []
    for r in range(R):
        ba, bb = r % 10, (3 * r + 1) % 10
        a = _draw(*bounds[ba], counts[r], g, power)
        b = _draw(*bounds[bb], counts[r], g, power)
        pairs.append((a, b, r))
    return _emit(pairs, lo, R, {f"type{i}": b for i, b in enumerate(bounds)})


def uniform_kg(num_nodes: int, num_edges: int, num_relations: int, seed: int = 0) -> KG:
    """The ``torch.randint`` graph of the reference's self-tests (src/models/rgcn.py:427-444)."""
    g = torch.Generator().manual_seed(seed)
    ei = torch.randint(0, num_nodes, (2, num_edges), generator=g, dtype=torch.int64)
    et = torch.randint(0, num_relations, (num_edges,), generator=g, dtype=torch.int64)
    return KG(ei, et, num_nodes, num_relations, {"all": (0, num_nodes)})


def scaled_kg(num_nodes: int, num_directed_edges: int, num_relations: int, seed: int = 42,
              power: float = 2.5, device="cpu") -> KG:
    """Large single-block graph (cfg5 family: 10 M nodes / 400 M edges / 30 relations),
    generated on ``device`` so the 400 M-edge case never touches host memory."""
    g = torch.Generator(device=device).manual_seed(seed)
    und = num_directed_edges // 2
    u = torch.rand(und, generator=g, device=device)
    a = torch.floor(num_nodes * u.pow(power)).clamp_(max=num_nodes - 1).to(torch.int64)
    u = torch.rand(und, generator=g, device=device)
    b = torch.floor(num_nodes * u.pow(power)).clamp_(max=num_nodes - 1).to(torch.int64)
    # hubs sit at low ids; a fixed odd-multiplier permutation spreads them over the id range
    mult = 2_654_435_761 % num_nodes
    while _gcd(mult, num_nodes) != 1:
        mult += 1
    a = (a * mult) % num_nodes
    b = (b * mult) % num_nodes
    r = torch.randint(0, num_relations, (und,), generator=g, device=device, dtype=torch.int64)
    ei = torch.stack([torch.stack([a, b], 1).reshape(-1), torch.stack([b, a], 1).reshape(-1)], 0)
    et = r.repeat_interleave(2)
    return KG(ei.contiguous(), et.contiguous(), num_nodes, num_relations, {"all": (0, num_nodes)})


def _gcd(a: int, b: int) -> int:
    while b:
        a, b = b, a % b
    return a


def link_batch(kg: KG, num_pos: int = 1024, seed: int = 42):
    """One training batch as built by reference src/train.py:276-288: ``num_pos`` graph edges as
    positives + one corruption each (head XOR tail replaced by a uniform node, train.py:59-97),
    labels 1/0.  Precomputed on the CPU generator so every implementation sees the same batch."""
    g = torch.Generator().manual_seed(seed + 1)
    E = kg.num_edges
    sel = torch.randperm(E, generator=g)[:num_pos]
    ph, pt, pr = kg.edge_index[0, sel].cpu(), kg.edge_index[1, sel].cpu(), kg.edge_type[sel].cpu()
    corrupt_head = torch.rand(num_pos, generator=g) < 0.5
    rnd = torch.randint(0, kg.num_nodes, (num_pos,), generator=g, dtype=torch.int64)
    nh = torch.where(corrupt_head, rnd, ph)
    nt = torch.where(~corrupt_head, rnd, pt)
    heads = torch.cat([ph, nh])
    tails = torch.cat([pt, nt])
    rels = torch.cat([pr, pr])
    labels = torch.cat([torch.ones(num_pos), torch.zeros(num_pos)])
    return heads, tails, rels, labels
