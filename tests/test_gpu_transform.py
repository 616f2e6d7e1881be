"""GPU tests of the tcgen05 relational-transform kernels (bf16-plane operands fed by TMA) against fp64 torch matmuls.

Tolerances: mode "fp32" (bf16 hi/lo planes, 3 products) — 1e-4 of the output scale, inside BASELINE.json's
rtol 1e-4; mode "bf16" — 2e-2."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

SHAPES = [  # (rows, K1, K2, d_out)
    (300, 48, 16, 32),          # golden-test dims: R=3, d_in=16, d_out=32
    (140, 96, 16, 24),          # d_out not a multiple of 32
    (100, 36, 12, 20),          # nothing a multiple of 8: padded plane strides
    (1000, 192, 64, 128),       # cfg1 layer 1
    (5000, 384, 128, 128),      # cfg1 layer 2
    (4097, 768, 256, 256),      # cfg2 layer 2, ragged row count
    (130, 64, 0, 512),          # two N tiles, single source
]
TOL = {"fp32": 1e-4, "bf16": 2e-2}


def _err(got, want):
    return float((got.double() - want).abs().max() / (want.abs().max() + 1e-30))


@pytest.fixture(scope="module")
def ops(lib_built):
    from primekg_rgcn_linkprediction_b200 import ops
    return ops


def _planes(ops, mats, mode, relu_mask=None, colsum=False):
    """fp32 matrices (concatenated along columns) -> bf16 planes."""
    n = mats[0].size(0)
    total = sum(m.size(1) for m in mats)
    P = ops.alloc_planes(n, total, mode, mats[0].device)
    c0, part = 0, None
    for m in mats:
        part = ops.split_planes(m, P, col0=c0, relu_mask=relu_mask, colsum=colsum)
        c0 += m.size(1)
    return P, part


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_split_planes(ops, mode):
    torch.manual_seed(3)
    x = torch.randn(777, 100, device=DEV) * 3
    m = torch.randn(777, 100, device=DEV)
    P, part = _planes(ops, [x], mode, relu_mask=m, colsum=True)
    want = x * (m > 0)
    hi = P[0].float()
    assert torch.equal(P[0], want.to(torch.bfloat16))
    if mode == "fp32":
        torch.testing.assert_close(hi + P[1].float(), want, rtol=2 ** -16, atol=1e-30)
    torch.testing.assert_close(part.double().sum(0), want.double().sum(0), rtol=1e-6, atol=1e-4)


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("relu", [False, True])
def test_transform_fwd(ops, shape, mode, relu):
    n, K1, K2, N = shape
    torch.manual_seed(0)
    A1 = torch.randn(n, K1, device=DEV)
    A2 = torch.randn(n, K2, device=DEV) if K2 else None
    W1 = torch.randn(K1, N, device=DEV) / (K1 + K2) ** 0.5
    W2 = torch.randn(K2, N, device=DEV) / (K1 + K2) ** 0.5 if K2 else None
    b = torch.randn(N, device=DEV)
    A, _ = _planes(ops, [A1] + ([A2] if K2 else []), mode)
    out = ops.transform_fwd(A, K1, K2, W1, W2, b, relu, mode)
    want = A1.double() @ W1.double() + b.double()
    if K2:
        want = want + A2.double() @ W2.double()
    if relu:
        want = want.clamp(min=0)
    assert _err(out, want) < TOL[mode], (shape, mode, _err(out, want))
    out2 = ops.transform_fwd(A, K1, K2, W1, W2, b, relu, mode)
    assert torch.equal(out, out2)


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("masked", [False, True])
def test_transform_dgrad(ops, shape, mode, masked):
    n, K1, K2, N = shape
    torch.manual_seed(1)
    gO = torch.randn(n, N, device=DEV)
    ro = torch.randn(n, N, device=DEV).clamp(min=0) if masked else None
    W1 = torch.randn(K1, N, device=DEV) / N ** 0.5
    W2 = torch.randn(K2, N, device=DEV) / N ** 0.5 if K2 else None
    G, _ = _planes(ops, [gO], mode, relu_mask=ro)
    gA = ops.transform_dgrad(G, N, W1, W2, mode)
    g = gO.double() * (ro > 0) if masked else gO.double()
    W = torch.cat([W1, W2], 0) if K2 else W1
    want = g @ W.double().t()
    assert gA.shape == (n, K1 + K2)
    assert _err(gA, want) < TOL[mode], (shape, mode, _err(gA, want))


@pytest.mark.parametrize("shape", SHAPES + [(30926, 768, 256, 256)])
@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("masked", [False, True])
def test_transform_wgrad(ops, shape, mode, masked):
    n, K1, K2, N = shape
    torch.manual_seed(2)
    A1 = torch.randn(n, K1, device=DEV)
    A2 = torch.randn(n, K2, device=DEV) if K2 else None
    gO = torch.randn(n, N, device=DEV)
    ro = torch.randn(n, N, device=DEV).clamp(min=0) if masked else None
    A, _ = _planes(ops, [A1] + ([A2] if K2 else []), mode)
    G, part = _planes(ops, [gO], mode, relu_mask=ro, colsum=True)
    gW1, gW2, gb = ops.transform_wgrad(A, K1, K2, G, N, part, mode)
    g = gO.double() * (ro > 0) if masked else gO.double()
    assert _err(gW1, A1.double().t() @ g) < TOL[mode], (shape, mode, "gW1", _err(gW1, A1.double().t() @ g))
    if K2:
        assert _err(gW2, A2.double().t() @ g) < TOL[mode], (shape, mode, "gW2")
    assert _err(gb, g.sum(0)) < 1e-5, (shape, mode, "gbias", _err(gb, g.sum(0)))
    again = ops.transform_wgrad(A, K1, K2, G, N, part, mode)
    assert torch.equal(gW1, again[0]) and torch.equal(gb, again[2])          # deterministic split-K


# ---- weights converted once per layer call (rgcn_prepare_weights): forward reads them MN-major, dgrad K-major ----------
@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_prepared_weights_forward_and_dgrad(ops, shape, mode):
    """Same operands, same K order, same three products: the MN-major-B kernel must reproduce the K-major-B kernel's
    bits, whole and in row chunks (the chunked layer forward), and match the fp64 product."""
    n, K1, K2, N = shape
    torch.manual_seed(11)
    mats = [torch.randn(n, K1, device=DEV)] + ([torch.randn(n, K2, device=DEV)] if K2 else [])
    W1 = torch.randn(K1, N, device=DEV) * 0.3
    W2 = torch.randn(K2, N, device=DEV) * 0.3 if K2 else None
    bias = torch.randn(N, device=DEV)
    P, _ = _planes(ops, mats, mode)
    old = ops.transform_fwd(P, K1, K2, W1, W2, bias, True, mode)
    wp = ops.prepare_weights(W1, W2, mode)
    new = ops.transform_fwd_w(P, K1 + K2, wp, N, bias, True, mode)
    assert torch.equal(new, old)
    want = torch.relu(torch.cat(mats, 1).double() @ torch.cat([W1] + ([W2] if K2 else []), 0).double() + bias.double())
    assert _err(new, want) < TOL[mode]
    # row chunks into one output buffer
    out = torch.full_like(old, float("nan"))
    step = 128 * max(1, (n // 3) // 128) if n > 256 else n
    for r0 in range(0, n, step):
        r1 = min(r0 + step, n)
        Pc = (P[0][r0:r1], None if P[1] is None else P[1][r0:r1])
        ops.transform_fwd_w(Pc, K1 + K2, wp, N, bias, True, mode, row_offset=r0, out=out[r0:r1])
    assert torch.equal(out, old)
    # dgrad through the same planes
    G = torch.randn(n, N, device=DEV)
    Gp, _ = _planes(ops, [G], mode)
    g_old = ops.transform_dgrad(Gp, N, W1, W2, mode)
    g_new = ops.transform_dgrad(Gp, N, W1, W2, mode, w_planes=wp)
    assert torch.equal(g_new, g_old)


def test_chunked_dropout_mask_is_one_consistent_mask(ops):
    """Row-chunked calls hash the GLOBAL element index: the chunks' masks are the slices of the whole call's mask."""
    torch.manual_seed(2)
    n, K, N = 1024, 64, 128
    A = torch.zeros(n, K, device=DEV)
    P, _ = _planes(ops, [A], "bf16")
    W = torch.zeros(K, N, device=DEV)
    bias = torch.ones(N, device=DEV)
    ctr = ops.dropout_counter(A.device)
    wp = ops.prepare_weights(W, None, "bf16")
    whole = ops.transform_fwd_w(P, K, wp, N, bias, True, "bf16", 0.5, 77, ctr)
    parts = torch.empty_like(whole)
    for r0 in range(0, n, 256):
        ops.transform_fwd_w((P[0][r0:r0 + 256], None), K, wp, N, bias, True, "bf16", 0.5, 77, ctr, row_offset=r0,
                            out=parts[r0:r0 + 256])
    assert torch.equal(parts, whole)
    assert 0.45 < float((whole > 0).float().mean()) < 0.55
