"""Listed-rows forward of the last encoder layer (csrc/aggregate.cu LIST walk, rgcn_transform_fwd_w_rows,
rgcn_layer_fwd with rows != NULL): the training step of the reference reads only ``node_embeddings[head]`` /
``[tail]`` of the encoder output (src/models/rgcn.py:325-326), so the last RGCNConv computes those rows alone.
The listed rows must carry the SAME BITS as the dense layer (same walk order per row, same K order per output
row), and the step's gradients must agree with the dense formulation.
"""
import pytest
import torch
import torch.nn.functional as F

from test_gpu_parity import DEV, graphs, _product_model, _close_by_scale

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pkg(lib_built):
    return lib_built


def _layer_inputs(name, d_in, d_out, seed=5):
    ei, et, N, R = graphs()[name]
    gen = torch.Generator().manual_seed(seed)
    x = torch.randn(N, d_in, generator=gen).to(DEV)
    W = (torch.randn(R * d_in, d_out, generator=gen) / d_in ** 0.5).to(DEV)
    root = (torch.randn(d_in, d_out, generator=gen) / d_in ** 0.5).to(DEV)
    bias = torch.randn(d_out, generator=gen).to(DEV)
    return ei, et, N, R, x, W, root, bias, gen


def _pairs(ei, N, n, gen, with_hub=True):
    head = torch.randint(0, N, (n,), generator=gen)
    tail = torch.randint(0, N, (n,), generator=gen)
    if n > 8:
        head[1::5] = head[0]                                            # duplicates
        tail[2] = head[3]
    if with_hub and ei.numel():
        tail[0] = int(torch.bincount(ei[1], minlength=N).argmax())      # the biggest hub is listed
    return head.to(DEV), tail.to(DEV)


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("name,d_in,d_out,n", [("uniform_r30", 64, 128, 300), ("primekg_100k", 256, 256, 2048),
                                               ("val_fixture", 64, 64, 64), ("ragged", 8, 8, 3), ("one_relation", 128, 64, 50)])
def test_listed_layer_forward_equals_dense_rows(pkg, name, d_in, d_out, n, mode):
    from primekg_rgcn_linkprediction_b200 import ops
    ei, et, N, R, x, W, root, bias, gen = _layer_inputs(name, d_in, d_out)
    g = pkg.RelGraph.from_edges(ei.to(DEV), et.to(DEV), N, R)
    head, tail = _pairs(ei, N, n, gen)
    rows, slot = ops.rows_list_build(head, tail, N)
    assert torch.equal(rows, torch.cat([head, tail]))
    m_c = (2 * n + 127) // 128 * 128
    want_slot = torch.full((N,), m_c, dtype=torch.int32)
    for c in range(2 * n - 1, -1, -1):
        want_slot[int(rows[c])] = c
    assert torch.equal(slot.cpu(), want_slot)
    x16 = ops.to_bf16(x) if (mode == "bf16" and d_in % 8 == 0) else None
    out_d, A_d, _ = ops.layer_fwd(g, x, x, W, root, bias, False, mode, x_bf16=x16)
    out_l, A_l, _ = ops.layer_fwd(g, x, x, W, root, bias, False, mode, x_bf16=x16, rows=rows, slot=slot)
    assert A_l[0].shape[0] == m_c and out_l.shape == out_d.shape
    K = (R + 1) * d_in
    for pl, pd in zip(A_l, A_d):
        if pl is None:
            assert pd is None
            continue
        first = slot[rows].to(torch.int64) == torch.arange(2 * n, device=DEV)    # a row lives at its FIRST position
        assert torch.equal(pl[:2 * n, :K][first], pd[rows][:, :K][first])        # compact planes = the dense planes' rows
        assert not pl[:2 * n, :K][~first].any()                                   # later duplicates are zero rows
        assert not pl[2 * n:, :K].any()                                           # and so is the padding
    assert torch.equal(out_l[rows], out_d[rows])                        # same bits at every listed row


def test_listed_forward_without_hubs(pkg):
    """A list that avoids the hub rows skips every hub chunk; the slot map is mandatory."""
    from primekg_rgcn_linkprediction_b200 import ops
    ei, et, N, R, x, W, root, bias, gen = _layer_inputs("primekg_100k", 64, 64)
    g = pkg.RelGraph.from_edges(ei.to(DEV), et.to(DEV), N, R)
    deg = torch.bincount(ei[1], minlength=N)
    light = torch.nonzero(deg < 100).flatten()
    head = light[torch.randint(0, light.numel(), (500,), generator=gen)].to(DEV)
    tail = light[torch.randint(0, light.numel(), (500,), generator=gen)].to(DEV)
    rows, slot = ops.rows_list_build(head, tail, N)
    out_d, _, _ = ops.layer_fwd(g, x, x, W, root, bias, False, "fp32")
    out_a, _, _ = ops.layer_fwd(g, x, x, W, root, bias, False, "fp32", rows=rows, slot=slot)
    assert torch.equal(out_a[rows], out_d[rows])
    with pytest.raises(ValueError):
        ops.layer_fwd(g, x, x, W, root, bias, False, "fp32", rows=rows, slot=None)


def test_rows_list_build_parks_bad_indices(pkg):
    from primekg_rgcn_linkprediction_b200 import ops
    head = torch.tensor([3, -1, 5, 99], device=DEV)
    tail = torch.tensor([5, 2, 100, 0], device=DEV)
    rows, slot = ops.rows_list_build(head, tail, 10)
    assert rows.tolist() == [3, 0, 5, 0, 5, 2, 0, 0]
    s = slot.tolist()
    assert s[3] == 0 and s[5] == 2 and s[2] == 5 and s[0] == 7 and s[1] == 128


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_listed_layer_backward_equals_dense(pkg, mode):
    """layer_bwd on the compact planes of the listed forward (a_compact) against the dense backward of the same gO."""
    from primekg_rgcn_linkprediction_b200 import ops
    ei, et, N, R, x, W, root, bias, gen = _layer_inputs("primekg_100k", 128, 128)
    g = pkg.RelGraph.from_edges(ei.to(DEV), et.to(DEV), N, R)
    head, tail = _pairs(ei, N, 1500, gen)
    rows, slot = ops.rows_list_build(head, tail, N)
    _, A_d, wp = ops.layer_fwd(g, x, x, W, root, bias, False, mode)
    _, A_l, _ = ops.layer_fwd(g, x, x, W, root, bias, False, mode, rows=rows, slot=slot)
    gO = torch.zeros(N, 128)
    uniq = torch.unique(rows.cpu())
    gO[uniq] = torch.randn(uniq.numel(), 128, generator=gen)
    gO = gO.to(DEV)
    dense = ops.layer_bwd(g, gO, None, 1.0, A_d, W, root, 128, mode, True, True, True, True)
    comp = ops.layer_bwd(g, gO, None, 1.0, A_l, W, root, 128, mode, True, True, True, True, rows=rows, slot=slot,
                         a_compact=True, w_planes=wp)
    assert torch.equal(comp[0], dense[0])                               # g_x bit for bit
    for a, b, what in zip(comp[2:], dense[2:], ("g_weight", "g_root", "g_bias")):
        _close_by_scale(a, b, what)


def _step(pkg, g, monkeypatch, fwd: bool, mode="fp32", fused_loss=False):
    from primekg_rgcn_linkprediction_b200 import rowsparse
    monkeypatch.setenv("PRIMEKG_RGCN_SPARSE_FWD", "1" if fwd else "0")
    monkeypatch.setenv("PRIMEKG_RGCN_SPARSE_BWD", "1")
    monkeypatch.setattr(rowsparse, "MAX_FRACTION", 1e9)
    rowsparse.clear()
    m = _product_model(pkg, g, mode)
    m.train()
    for mod in m.modules():                                             # same masks in both runs: no dropout
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    ei, et = g["edge_index"].to(DEV), g["edge_type"].to(DEV)
    h, t, r, y = (g[k].to(DEV) for k in ("heads", "tails", "rels", "labels"))
    if fused_loss:
        loss, scores, _ = m.link_loss(ei, et, h, t, r, y)
    else:
        scores = m(ei, et, h, t, r)
        loss = F.binary_cross_entropy_with_logits(scores, y)
    loss.backward()
    return scores.detach().clone(), {k: p.grad.clone() for k, p in m.named_parameters()}


@pytest.mark.parametrize("fused_loss", [False, True])
@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("fixture", ["small_full", "small_basis"])
def test_training_step_listed_forward_equals_dense_forward(pkg, monkeypatch, fixture, mode, fused_loss):
    from conftest import load_golden
    g = load_golden(fixture)
    s1, g1 = _step(pkg, g, monkeypatch, True, mode, fused_loss)
    s0, g0 = _step(pkg, g, monkeypatch, False, mode, fused_loss)
    assert torch.equal(s1, s0)                                          # the listed rows carry the dense layer's bits
    for k in g0:
        _close_by_scale(g1[k], g0[k], k, rtol=1e-4, atol=2e-5)


def test_eval_and_no_grad_calls_keep_the_dense_cached_encoding(pkg, monkeypatch):
    from conftest import load_golden
    g = load_golden("small_full")
    m = _product_model(pkg, g, "fp32")
    ei, et = g["edge_index"].to(DEV), g["edge_type"].to(DEV)
    h, t, r = (g[k].to(DEV) for k in ("heads", "tails", "rels"))
    s = m.predict(ei, et, h, t, r)
    emb = m.get_embeddings(ei, et)
    want = (emb[h] * m.decoder.relation_embeddings.weight[r] * emb[t]).sum(-1)
    torch.testing.assert_close(s, want, rtol=1e-5, atol=1e-6)
    assert m.encoder._eval_cache is not None and torch.isfinite(emb).all()


@pytest.mark.parametrize("mode", ["fp32"])
def test_marked_sources_walk_equals_plain_row_sparse_walk(pkg, monkeypatch, mode):
    """layer_bwd on a graph much larger than the row list: with the sources marked from the forward CSR first
    (rgcn_aggregate_bwd_rows_marked) the input gradient carries the same bits as without, and as the dense backward."""
    from primekg_rgcn_linkprediction_b200 import ops
    ei, et, N, R, x, W, root, bias, gen = _layer_inputs("primekg_100k", 64, 64)
    g = pkg.RelGraph.from_edges(ei.to(DEV), et.to(DEV), N, R)
    head, tail = _pairs(ei, N, 200, gen)                                # 400 listed rows of 100,000: ratio 250 (> 64)
    rows, slot = ops.rows_list_build(head, tail, N)
    _, A_d, wp = ops.layer_fwd(g, x, x, W, root, bias, False, mode)
    gO = torch.zeros(N, 64)
    uniq = torch.unique(rows.cpu())
    gO[uniq] = torch.randn(uniq.numel(), 64, generator=gen)
    gO = gO.to(DEV)
    dense = ops.layer_bwd(g, gO, None, 1.0, A_d, W, root, 64, mode, True, True, False, False)
    monkeypatch.setenv("RGCN_MARK_SOURCES", "1")
    marked = ops.layer_bwd(g, gO, None, 1.0, A_d, W, root, 64, mode, True, True, False, False, rows=rows, slot=slot)
    monkeypatch.setenv("RGCN_MARK_SOURCES", "0")
    plain = ops.layer_bwd(g, gO, None, 1.0, A_d, W, root, 64, mode, True, True, False, False, rows=rows, slot=slot)
    assert torch.equal(marked[0], plain[0]) and torch.equal(marked[0], dense[0])
