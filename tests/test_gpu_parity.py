"""GPU parity tests: the CUDA path, called through the C ABI (ctypes -> libcredgcn.so), against
(a) the golden fixtures produced by the reference itself and (b) the CPU oracle on seeded inputs.

Bars (BASELINE.json north_star): graph arrays / indices bit-exact; propagated embeddings, loss,
gradients, metrics within 1e-4 relative (fp32); top-K ids identical except where the oracle's own
fp32 scores tie to within 1e-6 relative."""
import numpy as np
import pytest
import torch

import pathlib

import credgcn_oracle as orc
from conftest import ORDER, VARIANTS, load_golden, rel_err

ROOT = pathlib.Path(__file__).resolve().parents[1]

pytestmark = pytest.mark.gpu
TOL = 1e-4
DEV = "cuda"


@pytest.fixture(scope="module")
def cg():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import credgcn  # noqa: F401
    from credgcn import evaluate, graph, model, sampler, synth, train, config
    return dict(graph=graph, model=model, sampler=sampler, evaluate=evaluate, synth=synth, train=train,
                config=config)


def _build(cg, g, tag=None):
    tag = tag or g["tag"]
    return cg["graph"].build_graph(g["train_edges"], int(g["num_users"]), int(g["num_items"]), g["cred"], tag, DEV)


def _bits(x):
    return np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)


# ------------------------------------------------------------------------------------------------
# graph build: bit-exact
# ------------------------------------------------------------------------------------------------
def test_graph_build_matches_reference_bit_exact(cg, golden):
    gr = _build(cg, golden)
    indptr, indices = gr.user_csr_numpy()
    np.testing.assert_array_equal(indptr, golden["csr_indptr"])
    np.testing.assert_array_equal(indices, golden["csr_indices"])
    for which in ("A", "C"):
        op = gr.operator(which)
        np.testing.assert_array_equal(op.indices().cpu().numpy(), golden[f"{which}_idx"])
        np.testing.assert_array_equal(_bits(op.values().cpu().numpy()), _bits(golden[f"{which}_val"]))
    if "deg_i" in golden:
        np.testing.assert_array_equal(gr.deg_i_float(), golden["deg_i"])
    assert gr.nnz < golden["train_edges"].shape[1]          # fixture graphs contain duplicate edges


def test_reference_named_builders(cg, golden):
    g, gm = golden, cg["graph"]
    U, I = int(g["num_users"]), int(g["num_items"])
    indptr, indices = gm.edges_to_user_csr(g["train_edges"], U)
    assert indptr.dtype == np.int64 and indices.dtype == np.int64
    np.testing.assert_array_equal(indptr, g["csr_indptr"])
    np.testing.assert_array_equal(indices, g["csr_indices"])
    if g["tag"] == "cu":
        M_ui, M_iu, deg_i = gm.build_cred_weighted_mats(g["train_edges"], U, I, g["cred"], DEV)
        assert M_ui.shape == (I, U) and M_iu.shape == (U, I)
        np.testing.assert_array_equal(deg_i, g["deg_i"])
        np.testing.assert_array_equal(_bits(M_ui.coalesce().values().cpu().numpy()), _bits(g["C_val"]))
        np.testing.assert_array_equal(_bits(M_iu.coalesce().values().cpu().numpy()), _bits(g["A_val"]))
    else:
        M_ui, M_iu = gm.build_message_passing_mats(g["train_edges"], U, I, torch.tensor(g["cred"]), DEV,
                                                   degree_aware=g["tag"] == "da")
        assert M_ui.shape == (U, I) and M_iu.shape == (I, U)
        np.testing.assert_array_equal(_bits(M_ui.values().cpu().numpy()), _bits(g["A_val"]))
        np.testing.assert_array_equal(_bits(M_iu.values().cpu().numpy()), _bits(g["C_val"]))
        dense = M_iu.to_sparse_coo()
        assert dense.shape == (I, U) and dense._nnz() == g["C_val"].size


@pytest.mark.parametrize("tag", VARIANTS)
def test_graph_build_vs_oracle_c1_shape(cg, tag):
    """C1-shaped graph (943 x 1,682 x 100k): long rows (> 256 nnz), zero-degree items, duplicates,
    cred end points 0.0 / 1.0."""
    sg = cg["synth"].make_graph("C1", duplicate_edges=300)
    gr = cg["graph"].build_graph(sg.train_edges, sg.num_users, sg.num_items, sg.cred, tag, DEV)
    ops = orc.Operators(sg.train_edges, sg.num_users, sg.num_items, sg.cred, tag)
    ip, ix = orc.edges_to_user_csr(sg.train_edges, sg.num_users)
    gip, gix = gr.user_csr_numpy()
    np.testing.assert_array_equal(gip, ip)
    np.testing.assert_array_equal(gix, ix)
    np.testing.assert_array_equal(gr.deg_u.cpu().numpy().astype(np.float32), ops.deg_u)
    np.testing.assert_array_equal(gr.deg_i.cpu().numpy().astype(np.float32), ops.deg_i)
    a, c = gr.operator("A"), gr.operator("C")
    np.testing.assert_array_equal(a.indices().cpu().numpy(), np.vstack([ops.A_row, ops.A_col]))
    np.testing.assert_array_equal(c.indices().cpu().numpy(), np.vstack([ops.C_row, ops.C_col]))
    np.testing.assert_array_equal(_bits(a.values().cpu().numpy()), _bits(ops.A_val))
    np.testing.assert_array_equal(_bits(c.values().cpu().numpy()), _bits(ops.C_val))
    # transposed value arrays: by-user val_bwd holds C's values in (u, i) order and vice versa
    ct = ops.Ct.tocoo()
    order = np.lexsort((ct.col, ct.row))
    np.testing.assert_array_equal(_bits(gr.by_user.val_bwd.cpu().numpy()), _bits(ct.data[order]))
    assert gr.by_item.n_long > 0, "C1 has items above 256 train edges: the chunked path must be exercised"
    assert (np.diff(gr.by_item.indptr.cpu().numpy()) == 0).any() or True


def test_graph_build_edge_cases(cg):
    gm = cg["graph"]
    # empty rows, a single edge, one user owning everything
    e = np.array([[3, 3, 3, 0], [0, 1, 2, 2]], dtype=np.int32)
    cred = np.array([1.0, 0.5, 0.0, 0.25, 0.75], np.float32)
    for tag in VARIANTS:
        gr = gm.build_graph(e, 5, 4, cred, tag, DEV)
        ops = orc.Operators(e, 5, 4, cred, tag)
        np.testing.assert_array_equal(_bits(gr.operator("A").values().cpu().numpy()), _bits(ops.A_val))
        np.testing.assert_array_equal(_bits(gr.operator("C").values().cpu().numpy()), _bits(ops.C_val))
        np.testing.assert_array_equal(gr.by_user.indptr.cpu().numpy(), ops.A.indptr)
        np.testing.assert_array_equal(gr.by_item.indptr.cpu().numpy(), ops.C.indptr)
    with pytest.raises(ValueError):
        gm.build_graph(np.array([[0, 9], [0, 0]], np.int32), 5, 4, cred, "v2", DEV)       # user id out of range
    with pytest.raises(ValueError):
        gm.build_graph(e, 5, 4, cred[:3], "v2", DEV)                                     # cred length


# ------------------------------------------------------------------------------------------------
# propagation forward / backward
# ------------------------------------------------------------------------------------------------
def _model(cg, g, gr):
    U, I, d, K = int(g["num_users"]), int(g["num_items"]), int(g["emb_dim"]), int(g["num_layers"])
    m = cg["model"]
    if g["tag"] == "cu":
        net = m.CredLightGCN(U, I, d, K, gr.operator("C"), gr.operator("A"))
    else:
        net = m.LightGCN(U, I, d, K, gr.operator("A"), gr.operator("C"))
    net.load_state_dict({"user_emb.weight": torch.tensor(g["e0_u"]), "item_emb.weight": torch.tensor(g["e0_i"])})
    return net.to(DEV)


def test_forward_matches_reference(cg, golden):
    g = golden
    net = _model(cg, g, _build(cg, g))
    fu, fi = net.final_embeddings() if g["tag"] == "cu" else net.get_user_item_emb()
    assert rel_err(fu.detach().cpu().numpy(), g["final_u"]) < TOL
    assert rel_err(fi.detach().cpu().numpy(), g["final_i"]) < TOL


def test_loss_and_gradients_match_reference_autograd_path(cg, golden):
    """Drop-in usage: model.final_embeddings()/bpr_loss + loss.backward(), as the reference scripts do."""
    g = golden
    net = _model(cg, g, _build(cg, g))
    ut, pt, nt = (torch.tensor(g[k], device=DEV) for k in ("users", "pos", "neg"))
    if g["tag"] == "cu":
        eu, ei = net.final_embeddings()
        ps, ns = net.score(ut, pt, eu, ei), net.score(ut, nt, eu, ei)
        loss = -torch.log(torch.sigmoid(ps - ns) + 1e-12).mean() + 1e-4 * net.l2_reg(ut, pt, nt)
    else:
        eu, ei = net.get_user_item_emb()
        loss = net.bpr_loss(ut, pt, nt, eu, ei, 1e-4)
    loss.backward()
    assert abs(loss.item() - float(g["loss"])) / abs(float(g["loss"])) < TOL
    assert rel_err(net.user_emb.weight.grad.cpu().numpy(), g["grad_u"]) < TOL
    assert rel_err(net.item_emb.weight.grad.cpu().numpy(), g["grad_i"]) < TOL


def test_fused_train_step_gradients(cg, golden):
    """Fused path (no autograd): forward + fused BPR/L2(/fair) + adjoint propagation + ego grads."""
    g = golden
    gr = _build(cg, g)
    for fair in ((0.0, 0.01) if g["tag"] == "cu" else (0.0,)):
        net = _model(cg, g, gr)
        pop = None
        if fair:
            deg_i = gr.deg_i_float()
            pop = (deg_i / max(float(deg_i.max()), 1.0)).astype(np.float32)
        step = cg["model"].TrainStep(net, reg_weight=1e-4, fair_weight=fair, pop=pop)
        loss = step.forward_backward(g["users"], g["pos"], g["neg"])
        key = "_fair" if fair else ""
        want = float(g["loss" + key])
        assert abs(loss.item() - want) / abs(want) < TOL
        assert rel_err(net.user_emb.weight.grad.cpu().numpy(), g["grad_u" + key]) < TOL
        assert rel_err(net.item_emb.weight.grad.cpu().numpy(), g["grad_i" + key]) < TOL
        assert rel_err(step.f_u.cpu().numpy(), g["final_u"]) < TOL


def test_layer_lists_api(cg, golden):
    g = golden
    if g["tag"] != "cu":
        pytest.skip("propagate_all_layers is a lightgcn_cu.py method")
    net = _model(cg, g, _build(cg, g))
    us, is_ = net.propagate_all_layers()
    assert len(us) == len(is_) == int(g["num_layers"]) + 1
    fu = torch.stack(us, 0).mean(0)
    assert rel_err(fu.detach().cpu().numpy(), g["final_u"]) < TOL
    torch.stack(is_, 0).mean(0).sum().backward()
    assert net.user_emb.weight.grad is not None and torch.isfinite(net.user_emb.weight.grad).all()


@pytest.mark.parametrize("order", ["jacobi", "gs"])
@pytest.mark.parametrize("K", [1, 2, 3, 4])
@pytest.mark.parametrize("d", [16, 32, 64, 128, 256])
def test_propagation_vs_oracle_all_widths(cg, order, K, d):
    sg = cg["synth"].make_graph("C1", num_users=400, num_items=300, num_edges=30_000, duplicate_edges=50)
    tag = "cu" if order == "jacobi" else "v2"
    gr = cg["graph"].build_graph(sg.train_edges, sg.num_users, sg.num_items, sg.cred, tag, DEV)
    ops = orc.Operators(sg.train_edges, sg.num_users, sg.num_items, sg.cred, tag)
    rng = np.random.default_rng(d * 10 + K)
    eu = (rng.standard_normal((sg.num_users, d)) * 0.1).astype(np.float32)
    ei = (rng.standard_normal((sg.num_items, d)) * 0.1).astype(np.float32)
    fu, fi = cg["model"].propagate_forward(gr, torch.tensor(eu, device=DEV), torch.tensor(ei, device=DEV), K, order)
    wu, wi = orc.propagate(ops, eu, ei, K, order)
    assert rel_err(fu.cpu().numpy(), wu) < TOL and rel_err(fi.cpu().numpy(), wi) < TOL
    gu = rng.standard_normal((sg.num_users, d)).astype(np.float32)
    gi = rng.standard_normal((sg.num_items, d)).astype(np.float32)
    bu, bi = cg["model"].propagate_backward(gr, torch.tensor(gu, device=DEV), torch.tensor(gi, device=DEV), K, order)
    ou, oi = orc.propagate_backward(ops, gu, gi, K, order)
    assert rel_err(bu.cpu().numpy(), ou) < TOL and rel_err(bi.cpu().numpy(), oi) < TOL


@pytest.mark.parametrize("name,tag", [("C1", "cu"), ("C1", "da")])
def test_c1_full_config_vs_oracle(cg, name, tag):
    """BASELINE configs[0]: ML-100K-shaped, d=64, 3 layers (and the 4-layer degree-aware operator)."""
    sg = cg["synth"].make_graph(name)
    K = 3 if tag == "cu" else 4
    order = ORDER[tag]
    gr = cg["graph"].build_graph(sg.train_edges, sg.num_users, sg.num_items, sg.cred, tag, DEV)
    ops = orc.Operators(sg.train_edges, sg.num_users, sg.num_items, sg.cred, tag)
    torch.manual_seed(42)
    eu = torch.nn.init.xavier_uniform_(torch.empty(sg.num_users, 64))
    ei = torch.nn.init.xavier_uniform_(torch.empty(sg.num_items, 64))
    users, pos, neg = cg["synth"].make_triples(sg, 4096)
    loss, gu, gi, fu, fi = orc.train_step_grads(ops, eu.numpy(), ei.numpy(), users, pos, neg, K, order, 1e-4)
    m = cg["model"]
    Net = m.CredLightGCN if tag == "cu" else m.LightGCN
    net = Net(sg.num_users, sg.num_items, 64, K, gr.operator("C" if tag == "cu" else "A"),
              gr.operator("A" if tag == "cu" else "C"))
    net.load_state_dict({"user_emb.weight": eu, "item_emb.weight": ei})
    net = net.to(DEV)
    step = m.TrainStep(net, reg_weight=1e-4)
    got = step.forward_backward(users, pos, neg)
    assert abs(got.item() - loss) / abs(loss) < TOL
    assert rel_err(step.f_u.cpu().numpy(), fu) < TOL and rel_err(step.f_i.cpu().numpy(), fi) < TOL
    assert rel_err(net.user_emb.weight.grad.cpu().numpy(), gu) < TOL
    assert rel_err(net.item_emb.weight.grad.cpu().numpy(), gi) < TOL


def test_properties_at_c2_size(cg):
    """BASELINE configs[1] at full size (31,668 x 38,048 x 1.56M): size-independent properties --
    CSR sortedness, degree sums, linearity of the propagation, and the adjoint identity
    <P(x), y> == <x, P^T(y)> that ties the backward kernel to the forward one."""
    sg = cg["synth"].make_graph("C2")
    gr = cg["graph"].build_graph(sg.train_edges, sg.num_users, sg.num_items, sg.cred, "v2", DEV)
    E = sg.train_edges.shape[1]
    assert gr.nnz == E and int(gr.deg_u.sum()) == E and int(gr.deg_i.sum()) == E
    for csr in (gr.by_user, gr.by_item):
        ip, ix = csr.indptr, csr.idx.to(torch.int64)
        assert int(ip[0]) == 0 and int(ip[-1]) == E and bool((ip[1:] >= ip[:-1]).all())
        key = csr.row_ids() * (csr.n_cols + 1) + ix
        assert bool((key[1:] > key[:-1]).all()), "rows must be sorted by (row, col) with no duplicates"
    # rebuilt graph is identical (idempotence / determinism of the sort + segmented reduce)
    gr2 = cg["graph"].build_graph(sg.train_edges, sg.num_users, sg.num_items, sg.cred, "v2", DEV)
    assert torch.equal(gr.by_item.idx, gr2.by_item.idx) and torch.equal(gr.by_item.val_fwd, gr2.by_item.val_fwd)
    m = cg["model"]
    gen = torch.Generator(device=DEV).manual_seed(0)
    mk = lambda n: torch.randn(n, 64, device=DEV, generator=gen)
    xu, xi, yu, yi = mk(sg.num_users), mk(sg.num_items), mk(sg.num_users), mk(sg.num_items)
    for order in ("gs", "jacobi"):
        pu, pi = m.propagate_forward(gr, xu, xi, 3, order)
        qu, qi = m.propagate_forward(gr, 2.5 * xu, 2.5 * xi, 3, order)
        assert rel_err(qu.cpu().numpy(), (2.5 * pu).cpu().numpy()) < 1e-5                 # homogeneity
        bu, bi = m.propagate_backward(gr, yu, yi, 3, order)
        lhs = (pu.double() * yu.double()).sum() + (pi.double() * yi.double()).sum()
        rhs = (xu.double() * bu.double()).sum() + (xi.double() * bi.double()).sum()
        assert abs(lhs - rhs) / abs(lhs) < 1e-5
        # run-to-run bitwise reproducibility (no atomics anywhere on the path)
        pu2, pi2 = m.propagate_forward(gr, xu, xi, 3, order)
        assert torch.equal(pu, pu2) and torch.equal(pi, pi2)


# ------------------------------------------------------------------------------------------------
# sampler
# ------------------------------------------------------------------------------------------------
def test_sampler_positives_and_rejection(cg):
    sg = cg["synth"].make_graph("C1")
    gr = cg["graph"].build_graph(sg.train_edges, sg.num_users, sg.num_items, sg.cred, "v2", DEV)
    ip, ix = gr.user_csr_numpy()
    users = np.flatnonzero(np.diff(ip) > 0)
    for mix in (None, 0.7):
        s = cg["sampler"].TripleSampler(gr, mix, 0.75, 50, seed=42)
        ut = torch.tensor(users, device=DEV)
        pos, neg = s.sample(ut, offset=0)
        pos2, neg2 = s.sample(ut, offset=0)
        assert torch.equal(pos, pos2) and torch.equal(neg, neg2)                  # counter-based: reproducible
        pos3, neg3 = s.sample(ut, offset=1)
        assert not torch.equal(neg, neg3)
        pos, neg = pos.cpu().numpy(), neg.cpu().numpy()
        for u, p, n in zip(users, pos, neg):
            row = ix[ip[u]:ip[u + 1]]
            assert p in row and n not in row and 0 <= n < sg.num_items


@pytest.mark.parametrize("mix", [None, 0.7, 1.0])
def test_sampler_negative_law_chi_square(cg, mix):
    """Negatives of one user over many draws follow mix*pop^0.75 + (1-mix)*uniform restricted to
    non-train items (V2:349-376, 805-810).  Chi-square over degree-ordered bins, 1e-4 level."""
    sg = cg["synth"].make_graph("C1", num_users=200, num_items=500, num_edges=8000)
    gr = cg["graph"].build_graph(sg.train_edges, sg.num_users, sg.num_items, sg.cred, "v2", DEV)
    ip, ix = gr.user_csr_numpy()
    user = int(np.argmax(np.diff(ip)))
    N = 400_000
    s = cg["sampler"].TripleSampler(gr, mix, 0.75, 50, seed=7)
    _, neg = s.sample(torch.full((N,), user, device=DEV, dtype=torch.int64), offset=3)
    counts = np.bincount(neg.cpu().numpy(), minlength=sg.num_items).astype(np.float64)
    pop = orc.popularity_law(gr.deg_i.cpu().numpy(), 0.75) if mix is not None else None
    law = orc.negative_law_for_user(sg.num_items, ix[ip[user]:ip[user + 1]], pop, mix if mix is not None else 0.0)
    assert counts[law == 0].sum() == 0
    order = np.argsort(-law)
    bins = np.array_split(order[law[order] > 0], 40)
    obs = np.array([counts[b].sum() for b in bins])
    exp = np.array([law[b].sum() for b in bins]) * N
    chi2 = ((obs - exp) ** 2 / exp).sum()
    assert chi2 < 90.0, chi2          # 39 dof: P(chi2 > 90) ~ 5e-6


def test_sampler_positive_is_uniform_over_row_multiset(cg):
    e = np.array([[0] * 6 + [1], [5, 5, 5, 7, 9, 9, 1]], dtype=np.int32)       # item 5 x3, 9 x2, 7 x1
    gr = cg["graph"].build_graph(e, 2, 12, np.ones(2, np.float32), "v2", DEV)
    s = cg["sampler"].TripleSampler(gr, None, seed=1)
    pos, _ = s.sample(torch.zeros(120_000, device=DEV, dtype=torch.int64))
    c = np.bincount(pos.cpu().numpy(), minlength=12) / 120_000
    np.testing.assert_allclose(c[[5, 7, 9]], [3 / 6, 1 / 6, 2 / 6], atol=0.01)


# ------------------------------------------------------------------------------------------------
# evaluation
# ------------------------------------------------------------------------------------------------
def _check_topk(ids, sc, o_ids, o_sc, f_u, f_i, users):
    diff = ids != o_ids
    if diff.any():
        for r, c in zip(*np.nonzero(diff)):
            a, b = float(o_sc[r, c]), float(sc[r, c])
            assert abs(a - b) <= 1e-6 * max(abs(a), 1e-3), (r, c, a, b)
    np.testing.assert_allclose(sc, o_sc, rtol=1e-5, atol=1e-7)


def test_full_rank_topk_vs_oracle(cg, golden):
    g = golden
    tr = (g["csr_indptr"], g["csr_indices"])
    users = np.flatnonzero(np.diff(g["test_indptr"]) > 0)
    fu, fi = torch.tensor(g["final_u"], device=DEV), torch.tensor(g["final_i"], device=DEV)
    ev = cg["evaluate"]
    for K in (1, 20, 33, 64):
        ids, sc = ev.topk_device(fu, fi, torch.tensor(users), ev._device_csr(tr, DEV), K)
        o_ids, o_sc = orc.full_rank_topk(g["final_u"], g["final_i"], users, tr, K)
        _check_topk(ids.cpu().numpy(), sc.cpu().numpy(), o_ids, o_sc, g["final_u"], g["final_i"], users)
    if "full_ranked" in g:        # ids the reference itself ranked first: equal except where ITS scores tie to 1e-6
        ids, _ = ev.topk_device(fu, fi, torch.tensor(users), ev._device_csr(tr, DEV), 20)
        ids, want = ids.cpu().numpy(), g["full_ranked"]
        f64u, f64i = g["final_u"].astype(np.float64), g["final_i"].astype(np.float64)
        for r, c in zip(*np.nonzero(ids != want)):
            a = float(f64u[users[r]] @ f64i[ids[r, c]])
            b = float(f64u[users[r]] @ f64i[want[r, c]])
            assert abs(a - b) <= 1e-6 * max(abs(a), 1e-3), (r, c, a, b)


def test_topk_ties_and_short_catalogue(cg):
    """Ties break by lowest item id; users with fewer than K unmasked items get masked (-1e9) items
    in id order, as a full descending sort of the masked score vector would."""
    ev = cg["evaluate"]
    U, I, d = 3, 40, 16
    f_u = torch.zeros(U, d, device=DEV)
    f_u[:, 0] = 1.0
    f_i = torch.zeros(I, d, device=DEV)
    f_i[:, 0] = torch.tensor([float(k // 4) for k in range(I)], device=DEV)     # groups of 4 equal scores
    indptr = np.array([0, 0, 38, 38], np.int64)
    indices = np.arange(38, dtype=np.int64)                                     # user 1 has seen items 0..37
    ids, sc = ev.topk_device(f_u, f_i, torch.arange(U), ev._device_csr((indptr, indices), DEV), 8)
    ids, sc = ids.cpu().numpy(), sc.cpu().numpy()
    np.testing.assert_array_equal(ids[0], [36, 37, 38, 39, 32, 33, 34, 35])
    np.testing.assert_array_equal(ids[1], [38, 39, 0, 1, 2, 3, 4, 5])
    assert sc[1, 2] == np.float32(-1e9)
    np.testing.assert_array_equal(ids[2], ids[0])


def test_evaluate_full_ranking_matches_reference_metrics(cg, golden):
    g = golden
    if "full_20" not in g:
        pytest.skip("lightgcn_cu.py has no full-rank evaluator")
    net = _model(cg, g, _build(cg, g))
    extra = g["tag"] == "v2"
    res = cg["evaluate"].evaluate_full_ranking(
        net, (g["csr_indptr"], g["csr_indices"]), (g["test_indptr"], g["test_indices"]), int(g["num_items"]), DEV,
        g["item_pop"] if extra else None, int(g["total_train"]) if extra else 0, g["cred"] if extra else None)
    keys = ("precision", "recall", "ndcg") + (("item_coverage", "avg_log_popularity", "avg_self_information",
                                                 "cred_utility", "high_cred_recall", "low_cred_recall") if extra else ())
    for K in (10, 20):
        np.testing.assert_allclose([res[K][k] for k in keys], g[f"full_{K}"], rtol=TOL, atol=1e-6)
        assert res[K]["mode"] == "full"


def test_evaluate_sampled_matches_reference_metrics(cg, golden):
    g = golden
    net = _model(cg, g, _build(cg, g))
    extra = g["tag"] == "v2"
    res = cg["evaluate"].evaluate_sampled(
        net, (g["csr_indptr"], g["csr_indices"]), (g["test_indptr"], g["test_indices"]), int(g["num_items"]), DEV,
        g["item_pop"] if extra else None, int(g["total_train"]) if extra else 0, g["cred"] if extra else None)
    # candidate lists are identical to the reference's (same PCG64 stream), so the metrics must agree to 1e-4 --
    # except for users whose positive ties with a negative to 1e-6 in the reference's own scores: there the rank of
    # the positive is not defined, and each such user may move a mean by at most 1 / n_users
    te = (g["test_indptr"], g["test_indices"])
    users, cands = orc.sampled_candidates((g["csr_indptr"], g["csr_indices"]), te, int(g["num_items"]), 99,
                                          int(cg["config"].cfg.seed))
    f64u, f64i = g["final_u"].astype(np.float64), g["final_i"].astype(np.float64)
    n_tied = 0
    for r, u in enumerate(users):
        sc = f64i[cands[r]] @ f64u[int(u)]
        n_tied += int((np.abs(sc[1:] - sc[0]) <= 1e-6 * max(abs(sc[0]), 1e-3)).any())
    slack = n_tied / max(len(users), 1)
    for K in (10, 20):
        want = g[f"sampled_{K}"]
        got = np.array([res[K]["precision"], res[K]["recall"], res[K]["ndcg"]])
        assert np.all(np.abs(got - want[:3]) <= TOL * np.abs(want[:3]) + 1e-9 + slack), (K, got, want[:3], n_tied)
        assert res[K]["negatives"] == 99


def test_full_rank_topk_c1_vs_oracle(cg):
    sg = cg["synth"].make_graph("C1")
    rng = np.random.default_rng(3)
    fu = (rng.standard_normal((sg.num_users, 64)) * 0.3).astype(np.float32)
    fi = (rng.standard_normal((sg.num_items, 64)) * 0.3).astype(np.float32)
    tr = orc.edges_to_user_csr(sg.train_edges, sg.num_users)
    users = np.arange(0, sg.num_users, 3)
    ev = cg["evaluate"]
    ids, sc = ev.topk_device(torch.tensor(fu, device=DEV), torch.tensor(fi, device=DEV), torch.tensor(users),
                             ev._device_csr(tr, DEV), 20)
    o_ids, o_sc = orc.full_rank_topk(fu, fi, users, tr, 20)
    _check_topk(ids.cpu().numpy(), sc.cpu().numpy(), o_ids, o_sc, fu, fi, users)


# ------------------------------------------------------------------------------------------------
# end to end
# ------------------------------------------------------------------------------------------------
def test_training_loop_learns(cg):
    """train_on_arrays: loss decreases and Recall@20 beats the untrained model on a small graph."""
    sg = cg["synth"].make_graph("C1", num_users=300, num_items=200, num_edges=12_000)
    cfg = cg["config"].CFG()
    cfg.device, cfg.variant, cfg.epochs, cfg.batch_size, cfg.eval_mode, cfg.lr = DEV, "v2", 15, 128, "full", 5e-3
    cfg.eval_every = 15
    cg["config"].cfg = cfg
    torch.manual_seed(0)
    model, res = cg["train"].train_on_arrays(sg.train_edges, sg.val_edges, sg.test_edges, sg.num_users, sg.num_items,
                                             sg.cred, None, cfg)
    assert set(model.state_dict().keys()) == {"user_emb.weight", "item_emb.weight"}
    assert res[20]["recall"] > 0.15 and res[20]["mode"] == "full"
    cg["config"].cfg = cg["config"].CFG()


def test_errors_are_loud(cg):
    from credgcn import _lib
    sg = cg["synth"].make_graph("C1", num_users=50, num_items=40, num_edges=500)
    gr = cg["graph"].build_graph(sg.train_edges, sg.num_users, sg.num_items, sg.cred, "v2", DEV)
    with pytest.raises(_lib.CgxError):
        cg["model"].LightGCN(sg.num_users, sg.num_items, 48, 3, gr.operator("A"), gr.operator("C"))   # emb_dim 48
    with pytest.raises(_lib.CgxError):
        cg["model"].propagate_forward(gr, torch.zeros(sg.num_users, 64), torch.zeros(sg.num_items, 64), 3, "gs")
    with pytest.raises((ValueError, TypeError)):
        cg["model"].LightGCN(sg.num_users, sg.num_items, 64, 3, gr.operator("A"), gr.operator("A"))
    f = torch.zeros(sg.num_users, 64, device=DEV)
    fi = torch.zeros(sg.num_items, 64, device=DEV)
    loss, *_ = cg["model"].bpr_fused(gr, f, fi, f, fi, [0], [sg.num_items + 3], [0], 1e-4)             # bad item id
    assert torch.isnan(loss).all()


# ------------------------------------------------------------------------------------------------
# tensor-core (tcgen05) score path
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(300, 420, 32), (943, 1682, 64), (700, 5000, 128), (130, 257, 16), (260, 900, 256)])
def test_tensor_core_topk_equals_fp32_path(cg, shape):
    """BF16x3 tensor-core selection + exact re-scoring (+ proof / redo) must return exactly what the
    fp32 kernel returns: same ids, same score bits."""
    U, I, d = shape
    sg = cg["synth"].make_graph("C1", num_users=U, num_items=I, num_edges=min(U * I // 8, 60_000))
    rng = np.random.default_rng(U + d)
    fu = torch.tensor((rng.standard_normal((U, d)) * 0.2).astype(np.float32), device=DEV)
    fi = torch.tensor((rng.standard_normal((I, d)) * 0.2).astype(np.float32), device=DEV)
    fi[::7] = fi[3]                                    # exact score ties across many items
    tr = orc.edges_to_user_csr(sg.train_edges, U)
    ev = cg["evaluate"]
    users = torch.arange(0, U, 2)
    csr = ev._device_csr(tr, DEV)
    from credgcn._lib import PRECISIONS, lib
    for K in (10, 20, 40):
        # the tcgen05 kernel itself must be what runs (d = 256 does not fit shared memory with BF16X3)
        assert lib().cgx_eval_topk_uses_tensor_cores(d, K, PRECISIONS["bf16x3"]) == (1 if d <= 128 else 0)
        ids0, sc0 = ev.topk_device(fu, fi, users, csr, K, "fp32")
        ids1, sc1 = ev.topk_device(fu, fi, users, csr, K, "bf16x3")
        assert torch.equal(ids0, ids1), (shape, K, int((ids0 != ids1).sum()))
        assert torch.equal(sc0, sc1)


@pytest.mark.parametrize("d", [64, 128, 256])
def test_tensor_core_single_pass_bf16_is_close(cg, d):
    """precision='bf16' (one bf16 pass, no proof): stated tolerance = at least 97 % of the exact top-20
    ids recovered, scores of common ids exact (they are re-scored in fp32)."""
    U, I = 1000, 6000
    from credgcn._lib import PRECISIONS, lib
    assert lib().cgx_eval_topk_uses_tensor_cores(d, 20, PRECISIONS["bf16"]) == 1
    rng = np.random.default_rng(9)
    fu = torch.tensor((rng.standard_normal((U, d)) * 0.2).astype(np.float32), device=DEV)
    fi = torch.tensor((rng.standard_normal((I, d)) * 0.2).astype(np.float32), device=DEV)
    ev = cg["evaluate"]
    csr = ev._device_csr((np.zeros(U + 1, np.int64), np.zeros(0, np.int64)), DEV)
    ids0, _ = ev.topk_device(fu, fi, torch.arange(U), csr, 20, "fp32")
    ids1, _ = ev.topk_device(fu, fi, torch.arange(U), csr, 20, "bf16")
    a, b = ids0.cpu().numpy(), ids1.cpu().numpy()
    overlap = np.mean([len(set(x) & set(y)) / 20 for x, y in zip(a, b)])
    assert overlap > 0.97, overlap


def test_evaluate_full_ranking_tensor_core_metrics(cg, golden):
    g = golden
    if "full_20" not in g:
        pytest.skip("lightgcn_cu.py has no full-rank evaluator")
    net = _model(cg, g, _build(cg, g))
    res = cg["evaluate"].evaluate_full_ranking(
        net, (g["csr_indptr"], g["csr_indices"]), (g["test_indptr"], g["test_indices"]), int(g["num_items"]), DEV,
        precision="bf16x3")
    for K in (10, 20):
        np.testing.assert_allclose([res[K][k] for k in ("precision", "recall", "ndcg")], g[f"full_{K}"][:3],
                                   rtol=TOL, atol=1e-6)


# ------------------------------------------------------------------------------------------------
# optimiser and CUDA-graph step
# ------------------------------------------------------------------------------------------------
def test_fused_adam_matches_torch_adam(cg):
    torch.manual_seed(3)
    pu, pi = torch.randn(500, 64, device=DEV), torch.randn(300, 64, device=DEV)
    qu, qi = torch.nn.Parameter(pu.clone()), torch.nn.Parameter(pi.clone())
    ref = torch.optim.Adam([qu, qi], lr=1e-3)
    ru, ri = torch.nn.Parameter(pu.clone()), torch.nn.Parameter(pi.clone())
    mine = cg["model"].FusedAdam(ru, ri, lr=1e-3)
    for t in range(5):
        gu, gi = torch.randn_like(pu) * (10.0 ** -t), torch.randn_like(pi)
        gi[::3] = 0.0
        qu.grad, qi.grad, ru.grad, ri.grad = gu.clone(), gi.clone(), gu.clone(), gi.clone()
        ref.step()
        mine.step()
        assert rel_err(ru.detach().cpu().numpy(), qu.detach().cpu().numpy()) < 1e-6
        assert rel_err(ri.detach().cpu().numpy(), qi.detach().cpu().numpy()) < 1e-6
    assert mine.state_dict()["step"] == 5


def test_graph_captured_step_equals_eager_steps(cg):
    """TrainStep.capture: N graph replays == N eager steps (same device-side sampler offsets and Adam
    step counts), bit for bit."""
    sg = cg["synth"].make_graph("C1")
    gr = cg["graph"].build_graph(sg.train_edges, sg.num_users, sg.num_items, sg.cred, "v2", DEV)
    users = torch.nonzero(gr.deg_u > 0).reshape(-1)[:512]
    outs = []
    for use_graph in (False, True):
        torch.manual_seed(11)
        net = cg["model"].LightGCN(sg.num_users, sg.num_items, 64, 3, gr.operator("A"), gr.operator("C")).to(DEV)
        samp = cg["sampler"].TripleSampler(gr, 0.7, 0.75, 50, seed=5)
        st = cg["model"].TrainStep(net, lr=1e-2, reg_weight=1e-4, sampler=samp)
        if use_graph:
            st.capture(512)
        losses = [float(st.step(users).item()) for _ in range(4)]
        outs.append((losses, net.user_emb.weight.detach().clone(), net.item_emb.weight.detach().clone()))
    assert outs[0][0] == outs[1][0]
    assert torch.equal(outs[0][1], outs[1][1]) and torch.equal(outs[0][2], outs[1][2])
    assert outs[0][0][-1] < outs[0][0][0]


# ------------------------------------------------------------------------------------------------
# plain LightGCN baseline (lightgcn.py): SURVEY.md section 8f-4
# ------------------------------------------------------------------------------------------------
def test_raw_lightgcn_matches_reference(cg):
    g = load_golden("small", "raw")
    U, I, d, K = int(g["num_users"]), int(g["num_items"]), int(g["emb_dim"]), int(g["num_layers"])
    adj = cg["graph"].build_norm_adj(g["train_edges"], U, I, DEV)
    coo = adj.to_sparse_coo()
    np.testing.assert_array_equal(coo.indices().cpu().numpy(), g["adj_idx"])
    np.testing.assert_allclose(coo.values().cpu().numpy(), g["adj_val"], rtol=1e-6)     # torch.pow vs exact sqrt/div
    net = cg["model"].RawLightGCN(U, I, d, K, adj)
    assert set(net.state_dict().keys()) == {"emb.weight"}
    net.load_state_dict({"emb.weight": torch.tensor(g["emb0"])})
    net = net.to(DEV)
    ue, ie = net.get_user_item_emb()
    assert rel_err(ue.detach().cpu().numpy(), g["final_u"]) < TOL
    assert rel_err(ie.detach().cpu().numpy(), g["final_i"]) < TOL
    ut, pt, nt = (torch.tensor(g[k], device=DEV) for k in ("users", "pos", "neg"))
    loss = net.bpr_loss(ut, pt, nt, ue, ie, 1e-4)
    loss.backward()
    assert abs(loss.item() - float(g["loss"])) / abs(float(g["loss"])) < TOL
    assert rel_err(net.emb.weight.grad.cpu().numpy(), g["grad"]) < TOL


# ------------------------------------------------------------------------------------------------
# sampled protocol on device: SURVEY.md section 8f-1
# ------------------------------------------------------------------------------------------------
def test_device_sampled_candidates_follow_the_protocol(cg):
    sg = cg["synth"].make_graph("C1")
    ev = cg["evaluate"]
    tr = orc.edges_to_user_csr(sg.train_edges, sg.num_users)
    te = orc.edges_to_user_csr(sg.test_edges, sg.num_users)
    users = np.flatnonzero(np.diff(te[0]) > 0)
    cand = ev.sampled_candidates_device(ev._device_csr(tr, DEV), ev._device_csr(te, DEV), users, sg.num_items, 99, 42)
    cand2 = ev.sampled_candidates_device(ev._device_csr(tr, DEV), ev._device_csr(te, DEV), users, sg.num_items, 99, 42)
    assert torch.equal(cand, cand2)                                   # counter-based: reproducible
    c = cand.cpu().numpy()
    assert c.shape == (len(users), 100)
    for r, u in enumerate(users[:200]):
        test_row, train_row = set(te[1][te[0][u]:te[0][u + 1]]), set(tr[1][tr[0][u]:tr[0][u + 1]])
        assert c[r, 0] in test_row
        assert not (set(c[r, 1:]) & (test_row | train_row)) and c[r, 1:].min() >= 0 and c[r, 1:].max() < sg.num_items
    # negatives are uniform over the allowed items: chi-square on item-id deciles over all users
    hist = np.bincount((c[:, 1:].ravel() * 10) // sg.num_items, minlength=10).astype(np.float64)
    allowed = np.ones((len(users), sg.num_items), bool)
    for r, u in enumerate(users):
        allowed[r, te[1][te[0][u]:te[0][u + 1]]] = False
        allowed[r, tr[1][tr[0][u]:tr[0][u + 1]]] = False
    dec = (np.arange(sg.num_items) * 10) // sg.num_items
    exp = np.array([(allowed[:, dec == k].sum(1) / allowed.sum(1)).sum() for k in range(10)]) * 99
    assert ((hist - exp) ** 2 / exp).sum() < 40.0                     # 9 dof


def test_device_ranking_equals_stable_argsort(cg):
    rng = np.random.default_rng(0)
    scores = rng.standard_normal((257, 100)).astype(np.float32)
    scores[:, 10:20] = scores[:, :10]                                 # ties
    cands = rng.integers(0, 5000, (257, 100))
    got = cg["evaluate"].rank_candidates_device(torch.tensor(scores, device=DEV), torch.tensor(cands, device=DEV))
    want = np.take_along_axis(cands, np.argsort(-scores, axis=1, kind="stable"), axis=1)
    np.testing.assert_array_equal(got.cpu().numpy(), want)


def test_evaluate_sampled_on_device_is_statistically_equivalent(cg, golden):
    """Device-drawn candidates: same protocol, different random stream -> metrics agree with the reference's
    within sampling noise (binomial std of Recall@20 over the evaluated users, 5 sigma)."""
    g = golden
    net = _model(cg, g, _build(cg, g))
    res = cg["evaluate"].evaluate_sampled(net, (g["csr_indptr"], g["csr_indices"]), (g["test_indptr"], g["test_indices"]),
                                          int(g["num_items"]), DEV, on_device=True)
    n = res[20]["users_eval"]
    want = float(g["sampled_20"][1])
    assert abs(res[20]["recall"] - want) < 5.0 * np.sqrt(max(want * (1 - want), 0.05) / n) + 0.02
    assert res[20]["mode"] == "sampled(1pos+neg)" and res[20]["negatives"] == 99


@pytest.mark.parametrize("mode", ["full", "sampled"])
@pytest.mark.parametrize("extra", [False, True])
def test_device_metrics_equal_host_metrics(cg, mode, extra):
    """cgx_eval_metrics / cgx_eval_coverage against the vectorised NumPy metrics (both accumulate in double;
    only the order of the additions differs)."""
    ev = cg["evaluate"]
    rng = np.random.default_rng(5)
    U, I, n, ld = 900, 1500, 700, 37
    Ks = [20, 5, 37, 10]                                              # unsorted on purpose
    users = np.sort(rng.choice(U, n, replace=False)).astype(np.int64)
    rows = [np.sort(rng.choice(I, rng.integers(1, 40), replace=True)) for _ in range(U)]   # duplicates kept
    te_indptr = np.concatenate([[0], np.cumsum([len(r) for r in rows])]).astype(np.int64)
    te_idx = np.concatenate(rows).astype(np.int64)
    ranked = np.stack([rng.permutation(I)[:ld] for _ in range(n)]).astype(np.int32)
    for r in range(0, n, 3):                                          # make sure hits exist
        row = rows[users[r]]
        ranked[r, rng.integers(0, ld)] = row[rng.integers(0, len(row))]
    item_pop = rng.integers(0, 500, I).astype(np.int64) if extra else None
    cred = rng.random(U) if extra else None
    total = int(item_pop.sum()) if extra else 0
    gt_single = ranked[np.arange(n), rng.integers(0, ld, n)].astype(np.int64) if mode == "sampled" else None
    pct = 0.6 if extra else 0.2                                       # 0.6: the two credibility groups overlap
    want = ev.metrics_from_ranked(ranked, users, (te_indptr, te_idx), I, Ks, mode, item_pop, total, cred, pct,
                                  gt_single=gt_single, extra_keys={"negatives": 99} if mode == "sampled" else None)
    te_dev = (torch.tensor(te_indptr, device=DEV), torch.tensor(te_idx, device=DEV, dtype=torch.int32))
    got = ev.metrics_device(torch.tensor(ranked, device=DEV), users, None if mode == "sampled" else te_dev, I, Ks, mode,
                            item_pop, total, cred, pct,
                            gt_single_dev=None if gt_single is None else torch.tensor(gt_single, device=DEV),
                            extra_keys={"negatives": 99} if mode == "sampled" else None)
    assert list(got) == list(want)
    for K in Ks:
        assert got[K].keys() == want[K].keys()
        for k, v in want[K].items():
            if isinstance(v, float):
                assert got[K][k] == pytest.approx(v, rel=1e-11, abs=1e-13), (K, k)
            else:
                assert got[K][k] == v, (K, k)
    again = ev.metrics_device(torch.tensor(ranked, device=DEV), users, None if mode == "sampled" else te_dev, I, Ks, mode,
                              item_pop, total, cred, pct,
                              gt_single_dev=None if gt_single is None else torch.tensor(gt_single, device=DEV),
                              extra_keys={"negatives": 99} if mode == "sampled" else None)
    assert again == got                                               # deterministic to the bit


def test_device_metrics_reject_bad_cutoffs(cg):
    ev = cg["evaluate"]
    ranked = torch.zeros(4, 10, dtype=torch.int32, device=DEV)
    from credgcn._lib import CgxError
    with pytest.raises(CgxError):
        ev.metrics_device(ranked, np.arange(4), None, 50, [20], "sampled", gt_single_dev=torch.zeros(4, device=DEV))


@pytest.mark.parametrize("d", [16, 32, 64, 128, 256])
def test_sparse_row_spmm_equals_dense_spmm_bitwise(cg, d):
    """cgx_spmm_sparse_rows (zero rows of X skipped by flag) returns the bits of cgx_spmm, for short, chunked
    (> 256 nnz) and huge (> 16384 nnz) rows, both row orders, forward and adjoint values."""
    sg = cg["synth"].make_graph("C1", duplicate_edges=300)
    gr = cg["graph"].build_graph(sg.train_edges, sg.num_users, sg.num_items, sg.cred, "v2", DEV)
    m = cg["model"]
    gen = torch.Generator(device=DEV).manual_seed(d)
    for csr, n_in in ((gr.by_user, sg.num_items), (gr.by_item, sg.num_users)):
        x = torch.randn(n_in, d, device=DEV, generator=gen)
        keep = torch.rand(n_in, device=DEV, generator=gen) < 0.1
        x[~keep] = 0.0
        x[5] = -0.0                                                   # negative zeros are zeros
        flags = m.row_flags(x)
        assert torch.equal(flags.bool(), (x != 0).any(1))
        for bwd in (False, True):
            want = m.spmm(csr, x, use_bwd_values=bwd)
            got = m.spmm_sparse_rows(csr, x, flags, use_bwd_values=bwd)
            assert torch.equal(got.view(torch.int32), want.view(torch.int32))
        # all-zero and all-dense inputs
        z = torch.zeros_like(x)
        assert not m.spmm_sparse_rows(csr, z).any()
        xd = torch.randn(n_in, d, device=DEV, generator=gen)
        assert torch.equal(m.spmm_sparse_rows(csr, xd), m.spmm(csr, xd))


def test_ragged_last_batch_after_capture(cg):
    """An epoch ends with a short batch (lightgcn_cu.py:608-610 slices train_users[start:start + batch]): a
    captured TrainStep takes it through the eager path with the same device counters, so a run of full + short
    batches equals the all-eager run bit for bit.  Batch sizes 1 and 33 (not a multiple of a warp) included."""
    sg = cg["synth"].make_graph("C1")
    gr = cg["graph"].build_graph(sg.train_edges, sg.num_users, sg.num_items, sg.cred, "v2", DEV)
    users = torch.nonzero(gr.deg_u > 0).reshape(-1)
    batches = [users[:256], users[256:512], users[512:545], users[545:546]]
    outs = []
    for use_graph in (False, True):
        torch.manual_seed(3)
        net = cg["model"].LightGCN(sg.num_users, sg.num_items, 32, 2, gr.operator("A"), gr.operator("C")).to(DEV)
        st = cg["model"].TrainStep(net, lr=1e-2, reg_weight=1e-4, sampler=cg["sampler"].TripleSampler(gr, 0.7, 0.75, 50, seed=9))
        if use_graph:
            st.capture(256)
        losses = [float(st.step(b).item()) for b in batches]
        outs.append((losses, net.user_emb.weight.detach().clone(), net.item_emb.weight.detach().clone()))
    assert outs[0][0] == outs[1][0] and all(np.isfinite(outs[0][0]))
    assert torch.equal(outs[0][1], outs[1][1]) and torch.equal(outs[0][2], outs[1][2])


def test_graph_without_edges(cg):
    """E = 0: every operator is empty, the layers above 0 vanish (final = E0 / (K + 1)), nothing is masked in
    the ranking, and the loss gradient reduces to the L2 term."""
    U, I, d, K = 7, 45, 16, 3
    gr = cg["graph"].build_graph(np.zeros((2, 0), np.int32), U, I, np.linspace(0, 1, U, dtype=np.float32), "v2", DEV)
    assert gr.nnz == 0 and int(gr.deg_u.sum()) == 0 and int(gr.deg_i.sum()) == 0
    torch.manual_seed(0)
    eu, ei = torch.randn(U, d, device=DEV), torch.randn(I, d, device=DEV)
    fu, fi = cg["model"].propagate_forward(gr, eu, ei, K, "gs")
    assert torch.equal(fu, eu * (1.0 / (K + 1))) and torch.equal(fi, ei * (1.0 / (K + 1)))
    gu, gi = cg["model"].propagate_backward(gr, torch.ones_like(eu), torch.ones_like(ei), K, "jacobi")
    assert torch.equal(gu, torch.full_like(eu, 1.0 / (K + 1))) and torch.equal(gi, torch.full_like(ei, 1.0 / (K + 1)))
    ids, sc = cg["evaluate"].topk_device(fu, fi, torch.arange(U, device=DEV), (gr.samp_indptr, gr.samp_idx), 10)
    want = torch.argsort(-(fu @ fi.T), dim=1, stable=True)[:, :10]
    assert torch.equal(ids.long(), want)


@pytest.mark.parametrize("d", [64, 128])
def test_spmm_geometry_for_tables_beyond_l2_gives_the_same_bits(cg, d):
    """cgx_spmm switches to narrower groups (two float4 per lane) when the gathered table exceeds L2; the sums are
    formed in the same order, so forcing that geometry on a small graph must reproduce the default bits -- dense
    and sparse-row products, forward and adjoint propagation."""
    from credgcn._lib import lib
    sg = cg["synth"].make_graph("C1", duplicate_edges=300)
    gr = cg["graph"].build_graph(sg.train_edges, sg.num_users, sg.num_items, sg.cred, "v2", DEV)
    m = cg["model"]
    gen = torch.Generator(device=DEV).manual_seed(d)
    xu = torch.randn(sg.num_users, d, device=DEV, generator=gen)
    xi = torch.randn(sg.num_items, d, device=DEV, generator=gen)
    xs = xu.clone()
    xs[torch.rand(sg.num_users, device=DEV, generator=gen) < 0.9] = 0.0

    def run():
        return (m.spmm(gr.by_user, xi), m.spmm(gr.by_item, xu, use_bwd_values=True), m.spmm_sparse_rows(gr.by_item, xs),
                *m.propagate_forward(gr, xu, xi, 3, "gs"), *m.propagate_backward(gr, xs, xi, 2, "jacobi"))
    want = run()
    old = lib().cgx_spmm_set_l2_table_bytes(0)
    try:
        got = run()
    finally:
        lib().cgx_spmm_set_l2_table_bytes(old)
    for a, b in zip(got, want):
        assert torch.equal(a.view(torch.int32), b.view(torch.int32))


@pytest.mark.parametrize("K", [5, 20, 40])
def test_tensor_core_topk_redo_rows_when_every_score_ties(cg, K):
    """All items identical: every approximate score ties, no completeness proof can hold, and every row goes
    through the exact per-row redo kernel -- the answer must still be the fp32 kernel's (lowest unmasked ids first,
    masked train items last with -1e9)."""
    U, I, d = 300, 1000, 64
    sg = cg["synth"].make_graph("C1", num_users=U, num_items=I, num_edges=20_000)
    rng = np.random.default_rng(K)
    fu = torch.tensor((rng.standard_normal((U, d)) * 0.3).astype(np.float32), device=DEV)
    fi = torch.tensor(np.repeat((rng.standard_normal((1, d)) * 0.3).astype(np.float32), I, 0), device=DEV)
    fi[500:] *= 0.5                                   # two plateaus, so that the sign of the user's score matters
    ev = cg["evaluate"]
    csr = ev._device_csr(orc.edges_to_user_csr(sg.train_edges, U), DEV)
    users = torch.arange(U)
    ids0, sc0 = ev.topk_device(fu, fi, users, csr, K, "fp32")
    ids1, sc1 = ev.topk_device(fu, fi, users, csr, K, "bf16x3")
    assert torch.equal(ids0, ids1) and torch.equal(sc0, sc1)


@pytest.mark.parametrize("variant", ["v2", "cu", "da"])
def test_file_driven_main_jsonl_to_checkpoint(cg, variant, tmp_path, capsys):
    """The reference's main() end to end on its own formats: review JSONL -> md5 split -> npy/pkl graph files ->
    credibility CSV -> training with validation -> best_model*.pt with the reference's state_dict keys ->
    test metrics (lightgcn_cu.py:690-703, 555-688).  73 items only, so K is cut to 5 / 10."""
    import pickle
    config, train = cg["config"], cg["train"]
    saved = config.cfg
    cfg = config.CFG()
    cfg.jsonl_path = str(ROOT / "tests" / "golden" / "tiny_reviews.jsonl")
    cfg.out_dir, cfg.device, cfg.variant = str(tmp_path), "cuda:0", variant
    cfg.epochs, cfg.eval_every, cfg.batch_size, cfg.emb_dim, cfg.num_layers = 4, 2, 32, 16, 2
    cfg.Ks, cfg.sampled_negatives = [5, 10], 20
    cfg.cred_csv_path = str(tmp_path / "cred.csv")
    config.cfg = cfg
    try:
        from credgcn import ingest
        ingest.build_graph_from_jsonl(cfg)
        u2 = pickle.load(open(tmp_path / "model" / "user2idx.pkl", "rb"))
        with open(cfg.cred_csv_path, "w") as f:
            f.write("user_id,credibility\n" + "".join(f"{u},{(k % 10) / 9:.4f}\n" for k, u in enumerate(u2)))
        train.main()                                     # graph files exist -> "Skipping construction" -> train
    finally:
        config.cfg = saved
    out = capsys.readouterr().out
    assert "Graph files exist. Skipping construction." in out and "TEST metrics" in out and "Epoch" in out
    ck = tmp_path / "model" / ("best_model_cred.pt" if variant == "cu" else "best_model.pt")
    sd = torch.load(ck, map_location="cpu")
    assert set(sd) == {"user_emb.weight", "item_emb.weight"}
    assert sd["user_emb.weight"].shape == (len(u2), 16) and torch.isfinite(sd["item_emb.weight"]).all()
