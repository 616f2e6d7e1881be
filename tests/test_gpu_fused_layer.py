"""The fused layer forward (csrc/fused_layer.cu: neighbourhood walk -> shared memory -> tcgen05) against the two-kernel
path it replaces (aggregate.cu -> transform.cu) and against the oracle's RGCNConv restatement.

Reference operator: RGCNConv loop path, call sites /root/reference/src/models/rgcn.py:123, :128.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def pkg(lib_built):
    return lib_built


def _graphs():
    from primekg_rgcn_linkprediction_b200 import synth
    yield "primekg_200k", synth.primekg_subgraph(200_000, seed=3)            # hubs, chunk-wise degree order
    yield "uniform_70k", synth.uniform_kg(70_000, 300_000, 5, seed=8)        # no row order: consecutive tiles
    yield "small_ragged", synth.uniform_kg(333, 4_000, 3, seed=1)            # one partial wave, empty segments
    kg = synth.uniform_kg(1_000, 30_000, 2, seed=4)                          # a hub row with 20,000 in-edges of one type
    kg.edge_index[1, :20_000] = 7
    kg.edge_type[:20_000] = 1
    yield "one_hub", kg


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("d_in,d_out,relu", [(64, 256, True), (256, 256, False), (128, 96, True), (256, 128, True)])
def test_fused_layer_forward_equals_two_kernel_path(pkg, mode, d_in, d_out, relu):
    """Same sums in the walk (left to right in CSR order; only rows cut between two producer warps are added piecewise),
    so the operand planes agree to fp32 rounding; the transform visits K in another order (column slices of 128), so the
    outputs agree to fp32 accumulation order; the dropout masks are the same hash."""
    from primekg_rgcn_linkprediction_b200 import ops
    for name, kg in _graphs():
        ei, et = kg.edge_index.to(DEV), kg.edge_type.to(DEV)
        g = pkg.RelGraph.from_edges(ei, et, kg.num_nodes, kg.num_relations)
        torch.manual_seed(5)
        R = kg.num_relations
        x = torch.randn(kg.num_nodes, d_in, device=DEV)
        W = torch.randn(R * d_in, d_out, device=DEV) * 0.1
        root = torch.randn(d_in, d_out, device=DEV) * 0.1
        bias = torch.randn(d_out, device=DEV)
        outs = []
        for schedule in (1, 3):
            ctr = ops.dropout_counter(x.device)
            drop = (0.5, 123, ctr) if relu else (0.0, 0, None)
            out, A, wp = ops.layer_fwd(g, x, x, W, root, bias, relu, mode, *drop, pipeline=schedule)
            torch.cuda.synchronize()
            outs.append((out, A[0], A[1]))
        (o1, h1, l1), (o3, h3, l3) = outs
        # the saved operand planes: the same sums; a row whose segment is cut between two producer warps is added piecewise
        a1 = h1.float() + (l1.float() if mode == "fp32" else 0)
        a3 = h3.float() + (l3.float() if mode == "fp32" else 0)
        a_scale = float(a1.abs().max())
        assert float((a1 - a3).abs().max()) <= (1e-5 if mode == "fp32" else 8e-3) * a_scale, name
        assert float((a1 != a3).float().mean()) < (0.2 if mode == "fp32" else 0.02), name
        scale = float(o1.abs().max())
        tol = (2e-5 if mode == "fp32" else 2e-2) * scale
        assert float((o1 - o3).abs().max()) <= tol, (name, float((o1 - o3).abs().max()), scale)
        if relu:
            # identical dropout masks: an element dropped by one path is dropped by the other (both exactly zero), up to
            # pre-activations within rounding distance of zero
            differ = ((o1 == 0) != (o3 == 0)).float().mean()
            assert float(differ) < 1e-4, name


@pytest.mark.parametrize("d_in,d_out", [(64, 128), (256, 256)])
def test_fused_layer_forward_matches_oracle(pkg, d_in, d_out):
    from oracle import rgcn_ref
    from primekg_rgcn_linkprediction_b200 import ops, synth
    kg = synth.primekg_subgraph(60_000, seed=11)
    ei, et = kg.edge_index.to(DEV), kg.edge_type.to(DEV)
    g = pkg.RelGraph.from_edges(ei, et, kg.num_nodes, kg.num_relations)
    torch.manual_seed(2)
    R = kg.num_relations
    x = torch.randn(kg.num_nodes, d_in)
    W = torch.randn(R, d_in, d_out) * 0.1
    root = torch.randn(d_in, d_out) * 0.1
    bias = torch.randn(d_out)
    want = rgcn_ref.rgcn_conv_ref(x, kg.edge_index, kg.edge_type, W, root, bias)
    out, _, _ = ops.layer_fwd(g, x.to(DEV), x.to(DEV), W.reshape(R * d_in, d_out).to(DEV), root.to(DEV), bias.to(DEV),
                              False, "fp32", pipeline=3)
    torch.testing.assert_close(out.cpu(), want, rtol=1e-4, atol=1e-4 * float(want.abs().max()))
