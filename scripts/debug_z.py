import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import primekg_rgcn_linkprediction_b200 as pkg
from primekg_rgcn_linkprediction_b200 import ops, synth
from primekg_rgcn_linkprediction_b200.graph import get_graph
DEV = "cuda:0"
kg = synth.primekg_subgraph(20_000, seed=9)
R, B = kg.num_relations, 2
ei, et = kg.edge_index.to(DEV), kg.edge_type.to(DEV)
N = kg.num_nodes
g = get_graph(ei, et, N, R)
print("hubs fwd/bwd", g.fwd.n_hubs, g.bwd.n_hubs, "chunks", g.fwd.n_chunks, g.bwd.n_chunks)
torch.manual_seed(0)
for d in (64, 256):
    G = torch.randn(N, d, device=DEV)
    comp = torch.randn(R, B, device=DEV)
    P = torch.randn(N, B * d, device=DEV)
    src, dst = ei[0], ei[1]
    cnt = torch.bincount(dst * R + et, minlength=N * R).clamp(min=1).double()
    w = 1.0 / cnt[dst * R + et]
    S = torch.zeros(N, R, d, dtype=torch.float64, device=DEV)
    S.index_put_((src, et), G.double()[dst] * w[:, None], accumulate=True)
    T_ref = torch.einsum("rb,nrd->nbd", comp.double(), S).reshape(N, B * d)
    gc_ref = torch.einsum("nrd,nbd->rb", S, P.double().view(N, B, d))
    T, gc = ops.aggregate_fwd(g, G, comp=comp, transposed=True, dot_p=P)
    print(d, "T rel err", float((T.double() - T_ref).norm() / T_ref.norm()), "max abs", float((T.double() - T_ref).abs().max()))
    print(d, "gc rel err", float((gc.double() - gc_ref).norm() / gc_ref.norm()))
    rows_bad = ((T.double() - T_ref).abs().amax(1) > 1e-3).nonzero().flatten()
    print("bad rows", rows_bad[:10].tolist(), rows_bad.numel())
    # forward orientation too
    x = torch.randn(N, d, device=DEV)
    Sf = torch.zeros(N, R, d, dtype=torch.float64, device=DEV)
    Sf.index_put_((dst, et), x.double()[src], accumulate=True)
    Hf = Sf / cnt.view(N, R, 1)
    Z_ref = torch.einsum("rb,nrd->nbd", comp.double(), Hf).reshape(N, B * d)
    Z = ops.aggregate_fwd(g, x, comp=comp)
    print(d, "Z rel err", float((Z.double() - Z_ref).norm() / Z_ref.norm()))
    planes = ops.alloc_planes(N, B * d, "fp32", DEV)
    ops.aggregate_fwd(g, G, comp=comp, transposed=True, planes=planes)
    Tp = planes[0].double() + planes[1].double()
    print(d, "T planes rel err", float((Tp - T_ref).norm() / T_ref.norm()))
