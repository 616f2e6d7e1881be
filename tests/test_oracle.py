"""CPU tests of the oracle itself: against the golden fixtures produced by the UNMODIFIED reference
model file (tests/golden/make_golden.py), a hand-computed micro graph, an independent dense-adjacency
formulation and fp64 gradcheck."""
import pytest
import torch
import torch.nn.functional as F

from conftest import load_golden, load_val_graph
from oracle import rgcn_ref as O


def _oracle_model(g, dropout=0.0):
    m = O.ModelRef(g["num_nodes"], g["num_relations"], g["embedding_dim"], g["hidden_dim"], dropout, 0.0,
                   g["num_bases"])
    m.load_state_dict(g["state_dict"], strict=True)
    return m


@pytest.mark.parametrize("name", ["small_full", "small_basis", "small_default_init"])
def test_oracle_matches_reference_goldens(name):
    g = load_golden(name)
    m = _oracle_model(g)
    m.train()
    loss, scores = O.train_step_ref(m, g["edge_index"], g["edge_type"], g["heads"], g["tails"], g["rels"], g["labels"])
    # same ops in the same order as the reference file => bit-identical on the same CPU build; allow
    # the last ulp for a different BLAS blocking on another host
    torch.testing.assert_close(scores, g["scores"], rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(loss, g["loss"], rtol=1e-6, atol=1e-7)
    for k, p in m.named_parameters():
        torch.testing.assert_close(p.grad, g["grads"][k], rtol=1e-5, atol=1e-7, msg=lambda s: f"{k}: {s}")
    m.eval()
    with torch.no_grad():
        emb = m.encoder(g["edge_index"], g["edge_type"])
        torch.testing.assert_close(emb, g["embeddings"], rtol=1e-6, atol=1e-6)
        all_t = m.decoder.score_all_tails(emb[g["heads"][:8]], g["rels"][:8], emb)
        torch.testing.assert_close(all_t, g["all_tail_scores"], rtol=1e-6, atol=1e-5)
        dec = m.decoder(emb[g["heads"]], emb[g["tails"]], g["rels"])
        torch.testing.assert_close(dec, g["decoder_scores"], rtol=1e-6, atol=1e-6)


def test_oracle_state_dict_keys_are_the_references():
    g = load_golden("small_basis")
    m = O.ModelRef(g["num_nodes"], g["num_relations"], g["embedding_dim"], g["hidden_dim"], 0.0, 0.0, g["num_bases"])
    assert set(m.state_dict().keys()) == set(g["state_dict"].keys())
    assert "encoder.conv1.comp" in m.state_dict()
    g = load_golden("small_full")
    m = O.ModelRef(g["num_nodes"], g["num_relations"], g["embedding_dim"], g["hidden_dim"])
    assert set(m.state_dict().keys()) == set(g["state_dict"].keys())


def test_parameter_count_pins_layout():
    # reference results/results.json:29, guide/MODEL_ARCHITECTURE.md:149
    m = O.ModelRef(30926, 3)
    assert sum(p.numel() for p in m.parameters()) == 2_078_208


def test_micro_graph_exact():
    g = load_golden("micro")
    out = O.rgcn_conv_ref(g["x"], g["edge_index"], g["edge_type"], g["weight"], g["root"], g["bias"])
    assert torch.equal(out, g["expected"])


@pytest.mark.parametrize("bases", [None, 3])
def test_loop_form_equals_dense_adjacency_form(bases):
    torch.manual_seed(0)
    N, E, R, di, do = 23, 160, 4, 5, 7
    ei = torch.randint(0, N, (2, E))
    et = torch.randint(0, R, (E,))
    x = torch.randn(N, di, dtype=torch.float64)
    W = torch.randn(bases or R, di, do, dtype=torch.float64)
    comp = torch.randn(R, bases, dtype=torch.float64) if bases else None
    root = torch.randn(di, do, dtype=torch.float64)
    b = torch.randn(do, dtype=torch.float64)
    a = O.rgcn_conv_ref(x, ei, et, W, root, b, comp)
    d = O.rgcn_conv_dense_ref(x, ei, et, W, root, b, comp)
    torch.testing.assert_close(a, d, rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("bases", [None, 2])
def test_oracle_gradcheck_fp64(bases):
    torch.manual_seed(1)
    N, E, R, di, do = 9, 40, 3, 3, 4
    ei = torch.randint(0, N, (2, E))
    et = torch.randint(0, R, (E,))
    mk = lambda *s: torch.randn(*s, dtype=torch.float64, requires_grad=True)
    x, W, root, b = mk(N, di), mk(bases or R, di, do), mk(di, do), mk(do)
    comp = mk(R, bases) if bases else None
    if bases:
        fn = lambda x, W, root, b, comp: O.rgcn_conv_ref(x, ei, et, W, root, b, comp)
        assert torch.autograd.gradcheck(fn, (x, W, root, b, comp))
    else:
        fn = lambda x, W, root, b: O.rgcn_conv_ref(x, ei, et, W, root, b)
        assert torch.autograd.gradcheck(fn, (x, W, root, b))


def test_csr_oracle_properties():
    v = load_val_graph()
    ei, et, N, R = v["edge_index"], v["edge_type"], v["num_nodes"], v["num_relations"]
    rowptr, col, perm, rowptr_t, row_t, perm_t = O.csr_oracle(ei, et, N, R)
    E = ei.size(1)
    assert rowptr[-1] == E and rowptr_t[-1] == E
    assert torch.equal(torch.sort(perm).values, torch.arange(E))
    key = ei[1] * R + et
    assert torch.all(key[perm][1:] >= key[perm][:-1])                    # sorted by (dst, rel)
    same = key[perm][1:] == key[perm][:-1]
    assert torch.all(perm[1:][same] > perm[:-1][same])                   # stable: original order inside a key
    assert torch.equal(col, ei[0][perm]) and torch.equal(row_t, ei[1][perm_t])
    # only relation 0 exists in this fixture: relations 1, 2 are empty segments
    cnt = (rowptr[1:] - rowptr[:-1]).view(N, R)
    assert int(cnt[:, 1:].sum()) == 0 and int(cnt[:, 0].sum()) == E


def test_csr_oracle_rejects_out_of_range():
    ei = torch.tensor([[0, 5], [1, 0]])
    with pytest.raises(IndexError):
        O.csr_oracle(ei, torch.tensor([0, 0]), 5, 1)
    with pytest.raises(IndexError):
        O.csr_oracle(torch.tensor([[0], [1]]), torch.tensor([3]), 5, 3)


def test_oracle_reproduces_reference_on_real_fixture_graph():
    """val_graph.npz holds rows of the reference model's output on the reference's own shipped
    validation graph; the oracle, seeded the same way, must reproduce them (init order included)."""
    v = load_val_graph()
    torch.manual_seed(v["seed"])
    m = O.ModelRef(v["num_nodes"], v["num_relations"], 64, 128, 0.5, 0.1)
    with torch.no_grad():
        for k, p in m.named_parameters():
            if "conv" in k and not k.endswith("bias"):
                p.mul_(v["conv_scale"])
    m.eval()
    with torch.no_grad():
        emb = m.encoder(v["edge_index"], v["edge_type"])
    torch.testing.assert_close(emb[v["rows"]], v["emb_rows"], rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(emb.double().sum(0), v["emb_colsum"], rtol=1e-6, atol=1e-6)


def test_rank_oracle_brackets_argsort_rank():
    torch.manual_seed(3)
    s = torch.randn(16, 50)
    s[0, 3] = s[0, 7]                                   # an exact tie
    true = torch.randint(0, 50, (16,))
    true[0] = 3
    opt, pes = O.rank_of_true_tail_ref(s, true)
    for i in range(16):                                 # reference src/evaluate.py:266-276
        order = torch.argsort(s[i], descending=True)
        rank = int((order == true[i]).nonzero()[0]) + 1
        assert int(opt[i]) <= rank <= int(pes[i])


def test_negative_batch_oracle_layout_and_rates():
    """The restated sampler + batch assembly (reference src/train.py:59-97, :281-288): the properties the device sampler
    is tested for on the GPU."""
    g = torch.Generator().manual_seed(0)
    N, n, k = 1000, 4096, 2
    ph, pt, pr = torch.randint(0, N, (n,), generator=g), torch.randint(0, N, (n,), generator=g), torch.randint(0, 3, (n,), generator=g)
    h, t, r, y = O.negative_batch_ref(ph, pt, pr, N, k, generator=g)
    assert torch.equal(h[:n], ph) and torch.equal(t[:n], pt) and torch.equal(r, torch.cat([pr, pr.repeat_interleave(k)]))
    assert y.tolist() == [1.0] * n + [0.0] * (n * k)
    keep_h, keep_t = h[n:] == ph.repeat_interleave(k), t[n:] == pt.repeat_interleave(k)
    assert bool((keep_h | keep_t).all())
    assert 0.47 < float((~keep_h).float().mean()) < 0.53
    assert O.accuracy_count_ref(torch.tensor([2.0, -1.0, 0.5]), torch.tensor([1.0, 0.0, 0.0])) == 2
