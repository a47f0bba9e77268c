"""Pin the CPU oracle to outputs of the reference itself (tests/golden/*.npz, made by
tests/golden/make_golden.py from /root/reference).  Integer/index/weight arrays are
bit-exact; floating point is held to 1e-4 relative (BASELINE.json north_star)."""
import numpy as np
import pytest

import credgcn_oracle as orc
from conftest import rel_err

TOL = 1e-4


def _ops(g):
    return orc.Operators(g["train_edges"], int(g["num_users"]), int(g["num_items"]), g["cred"], g["tag"])


def test_user_csr_bit_exact(golden):
    indptr, indices = orc.edges_to_user_csr(golden["train_edges"], int(golden["num_users"]))
    assert indptr.dtype == np.int64 and indices.dtype == np.int64
    np.testing.assert_array_equal(indptr, golden["csr_indptr"])
    np.testing.assert_array_equal(indices, golden["csr_indices"])


def test_operators_bit_exact(golden):
    ops = _ops(golden)
    np.testing.assert_array_equal(np.vstack([ops.A_row, ops.A_col]), golden["A_idx"])
    np.testing.assert_array_equal(np.vstack([ops.C_row, ops.C_col]), golden["C_idx"])
    assert ops.A_val.dtype == np.float32
    np.testing.assert_array_equal(ops.A_val.view(np.uint32), golden["A_val"].view(np.uint32))
    np.testing.assert_array_equal(ops.C_val.view(np.uint32), golden["C_val"].view(np.uint32))
    if "deg_i" in golden:
        np.testing.assert_array_equal(ops.deg_i, golden["deg_i"])
    # the fixture graphs contain duplicate (u, i) pairs: coalesce must have merged some
    assert ops.A_val.size < golden["train_edges"].shape[1]


def test_propagation(golden):
    fu, fi = orc.propagate(_ops(golden), golden["e0_u"], golden["e0_i"], int(golden["num_layers"]), golden["order"])
    assert rel_err(fu, golden["final_u"]) < TOL
    assert rel_err(fi, golden["final_i"]) < TOL


def test_loss_and_gradients(golden):
    g = golden
    reg = 1e-4                                   # cfg.lambda_reg (CU:58) == cfg.reg (V2:46)
    loss, gu, gi, _, _ = orc.train_step_grads(_ops(g), g["e0_u"], g["e0_i"], g["users"], g["pos"], g["neg"],
                                              int(g["num_layers"]), g["order"], reg)
    assert abs(loss - float(g["loss"])) / abs(float(g["loss"])) < TOL
    assert rel_err(gu, g["grad_u"]) < TOL
    assert rel_err(gi, g["grad_i"]) < TOL
    if g["tag"] == "cu":                         # fairness term, CU:641 with lambda_fair=0.01
        ops = _ops(g)
        pop = (ops.deg_i / max(float(ops.deg_i.max()), 1.0)).astype(np.float32)
        loss, gu, gi, _, _ = orc.train_step_grads(ops, g["e0_u"], g["e0_i"], g["users"], g["pos"], g["neg"],
                                                  int(g["num_layers"]), g["order"], reg, 0.01, pop)
        assert abs(loss - float(g["loss_fair"])) / abs(float(g["loss_fair"])) < TOL
        assert rel_err(gu, g["grad_u_fair"]) < TOL
        assert rel_err(gi, g["grad_i_fair"]) < TOL


def test_torch_cpu_baseline_leg_matches(golden):
    g = golden
    b = orc.TorchCpuBaseline(_ops(g), g["e0_u"], g["e0_i"], int(g["num_layers"]), g["order"])
    loss = b.step(g["users"], g["pos"], g["neg"], 1e-4, optimize=False)
    assert abs(loss - float(g["loss"])) / abs(float(g["loss"])) < TOL
    assert rel_err(b.eu.grad.numpy(), g["grad_u"]) < TOL
    assert rel_err(b.ei.grad.numpy(), g["grad_i"]) < TOL


def test_sampler_same_stream_same_triples(golden):
    """Same PCG64 seed and draw order as the reference loop => identical triples."""
    g = golden
    indptr, indices = g["csr_indptr"], g["csr_indices"]
    rng = np.random.default_rng(42)
    train_users = np.where(np.diff(indptr) > 0)[0]
    rng.shuffle(train_users)
    np.testing.assert_array_equal(train_users, g["shuffled_users"])
    batch = train_users[: len(g["samp_users"])]
    pop_prob = None
    if "pop_prob" in g:
        pop_prob = orc.popularity_law(orc.degrees(g["train_edges"], int(g["num_users"]), int(g["num_items"]))[1], 0.75)
        np.testing.assert_array_equal(pop_prob, g["pop_prob"])
    u, p, n = orc.sample_batch(indptr, indices, batch, int(g["num_items"]), rng, pop_prob, 0.7, 50)
    np.testing.assert_array_equal(u, g["samp_users"])
    np.testing.assert_array_equal(p, g["samp_pos"])
    np.testing.assert_array_equal(n, g["samp_neg"])


def test_sampled_protocol(golden):
    g = golden
    tr = (g["csr_indptr"], g["csr_indices"])
    te = (g["test_indptr"], g["test_indices"])
    users, cands = orc.sampled_candidates(tr, te, int(g["num_items"]), 99, 42)
    ranked = orc.rank_candidates(g["final_u"], g["final_i"], users, cands)
    np.testing.assert_array_equal(np.sort(ranked[:, :20], 1).shape, g["sampled_ranked"].shape)
    # identical candidate sets; order may differ only where scores tie to fp32 rounding
    same = (ranked[:, :20] == g["sampled_ranked"]).mean()
    assert same > 0.995
    gt = [{int(c[0])} for c in cands]
    extra = g["tag"] == "v2"
    res = orc.metrics_from_topk(ranked, users, te, int(g["num_items"]), (10, 20),
                                g["item_pop"] if extra else None, int(g["total_train"]) if extra else 0,
                                g["cred"] if extra else None, mode="sampled", gt_override=gt)
    for K in (10, 20):
        want = g[f"sampled_{K}"]
        got = [res[K]["precision"], res[K]["recall"], res[K]["ndcg"]]
        np.testing.assert_allclose(got, want[:3], rtol=TOL, atol=1e-6)
        if extra:
            got = [res[K][k] for k in ("item_coverage", "avg_log_popularity", "avg_self_information",
                                       "cred_utility", "high_cred_recall", "low_cred_recall")]
            np.testing.assert_allclose(got, want[3:], rtol=TOL, atol=1e-6)


def test_full_ranking(golden):
    g = golden
    if "full_ranked" not in g:
        pytest.skip("lightgcn_cu.py has no full-rank evaluator (SURVEY.md section 1, L6)")
    tr = (g["csr_indptr"], g["csr_indices"])
    te = (g["test_indptr"], g["test_indices"])
    users = np.flatnonzero(np.diff(te[0]) > 0)
    ids, sc = orc.full_rank_topk(g["final_u"], g["final_i"], users, tr, 20)
    ref = g["full_ranked"]
    diff = ids != ref
    # ids equal except where the oracle's own fp32 scores are within 1e-6 relative of each other
    if diff.any():
        r, c = np.nonzero(diff)
        for rr, cc in zip(r, c):
            s_mine = sc[rr, cc]
            s_ref = float((g["final_u"][users[rr]] * g["final_i"][ref[rr, cc]]).sum())
            assert abs(s_mine - s_ref) <= 1e-6 * max(abs(s_mine), 1e-3), (rr, cc, s_mine, s_ref)
    extra = g["tag"] == "v2"
    res = orc.evaluate_full_ranking(g["final_u"], g["final_i"], tr, te, int(g["num_items"]), (10, 20),
                                    g["item_pop"] if extra else None, int(g["total_train"]) if extra else 0,
                                    g["cred"] if extra else None)
    for K in (10, 20):
        want = g[f"full_{K}"]
        got = [res[K]["precision"], res[K]["recall"], res[K]["ndcg"]]
        np.testing.assert_allclose(got, want[:3], rtol=TOL, atol=1e-6)
        if extra:
            got = [res[K][k] for k in ("item_coverage", "avg_log_popularity", "avg_self_information",
                                       "cred_utility", "high_cred_recall", "low_cred_recall")]
            np.testing.assert_allclose(got, want[3:], rtol=TOL, atol=1e-6)


def test_plain_lightgcn_is_the_unit_credibility_jacobi_case():
    """lightgcn.py (no credibility, one N x N operator) == Jacobi propagation with c_u = 1 on both
    blocks: pins the reduction the product uses for SURVEY.md section 8f-4 against the reference's outputs."""
    from conftest import load_golden
    g = load_golden("small", "raw")
    U, I, K = int(g["num_users"]), int(g["num_items"]), int(g["num_layers"])
    ops = orc.Operators(g["train_edges"], U, I, np.ones(U, np.float32), "cu")
    e0 = g["emb0"]
    fu, fi = orc.propagate(ops, e0[:U], e0[U:], K, "jacobi")
    assert rel_err(fu, g["final_u"]) < TOL and rel_err(fi, g["final_i"]) < TOL
    loss, gu, gi, _, _ = orc.train_step_grads(ops, e0[:U], e0[U:], g["users"], g["pos"], g["neg"], K, "jacobi", 1e-4)
    assert abs(loss - float(g["loss"])) / abs(float(g["loss"])) < TOL
    assert rel_err(np.concatenate([gu, gi]), g["grad"]) < TOL
