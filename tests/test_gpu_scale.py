"""GPU parity at the BASELINE sizes, against the CPU oracle (not against properties), through the C ABI.

    C2  31,668 x 38,048 x 1.56 M, V2 operator, 3 layers, d = 64         (Version-2/lighgcn_cu_pop.py)
    C3  52,643 x 91,599 x 2.98 M, degree-aware operator, 4 layers, d = 64 (Degree-Aware Message.py:349-403, 424-442)
    C4s a C4-shaped 1/16 subsample (625,000 x 250,000 x 12.5 M, d = 128): both gathered tables exceed the 96 MiB
        regime switch WITHOUT forcing it, the hottest item row has > 16,384 non-zeros (finishing kernel), the hot-row
        hints are active in the kernels that run.

Bars: graph arrays bit-exact; forward / loss / gradients of one injected 4096-triple batch <= 1e-4 relative;
top-20 ids of 1,000 users identical except where the oracle's own fp32 scores tie to 1e-6."""
import numpy as np
import pytest
import torch

import credgcn_oracle as orc
from conftest import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = 1e-4


@pytest.fixture(scope="module")
def cg():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from credgcn import _lib, evaluate, graph, model, synth
    return dict(lib=_lib, graph=graph, model=model, evaluate=evaluate, synth=synth)


def _bits(x):
    return np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)


def _triples(sg, batch, seed):
    """Injected (user, pos, neg): pos from the user's row; neg uniform (set membership is irrelevant to the arithmetic)."""
    rng = np.random.default_rng(seed)
    u, i = sg.train_edges[0].astype(np.int64), sg.train_edges[1].astype(np.int64)
    sel = rng.integers(0, u.size, size=batch)
    return u[sel], i[sel], rng.integers(0, sg.num_items, size=batch).astype(np.int64)


def _check_topk(ids, sc, o_ids, o_sc):
    diff = ids != o_ids
    for r, c in zip(*np.nonzero(diff)):          # ids may differ only where the oracle's own scores tie
        a, b = float(o_sc[r, c]), float(sc[r, c])
        assert abs(a - b) <= 1e-6 * max(abs(a), 1e-3), (r, c, a, b)
    np.testing.assert_allclose(sc, o_sc, rtol=1e-5, atol=1e-7)


def _graph_arrays_bit_exact(gr, ops, sg):
    indptr, indices = gr.user_csr_numpy()
    o_indptr, o_indices = orc.edges_to_user_csr(sg.train_edges, sg.num_users)
    np.testing.assert_array_equal(indptr, o_indptr)
    np.testing.assert_array_equal(indices, o_indices)
    np.testing.assert_array_equal(gr.deg_u.cpu().numpy(), ops.deg_u.astype(np.int32))
    np.testing.assert_array_equal(gr.deg_i.cpu().numpy(), ops.deg_i.astype(np.int32))
    assert gr.nnz == ops.A_val.size
    for csr, rows, cols, v_fwd, v_bwd_src in (
            (gr.by_user, ops.A_row, ops.A_col, ops.A_val, ops.Ct),
            (gr.by_item, ops.C_row, ops.C_col, ops.C_val, ops.At)):
        np.testing.assert_array_equal(csr.indptr.cpu().numpy(),
                                      np.concatenate([[0], np.cumsum(np.bincount(rows, minlength=csr.n_rows))]))
        np.testing.assert_array_equal(csr.idx.cpu().numpy(), cols.astype(np.int32))
        np.testing.assert_array_equal(_bits(csr.val_fwd.cpu().numpy()), _bits(v_fwd))
        t = v_bwd_src.tocsr()
        t.sort_indices()
        np.testing.assert_array_equal(_bits(csr.val_bwd.cpu().numpy()), _bits(t.data))


def _step_parity(cg, sg, gr, ops, d, K, order, batch=4096, n_eval=1000, check_graph=True):
    m, ev = cg["model"], cg["evaluate"]
    if check_graph:
        _graph_arrays_bit_exact(gr, ops, sg)
    torch.manual_seed(42)
    Net = m.CredLightGCN if order == "jacobi" else m.LightGCN
    opsv = (gr.operator("C"), gr.operator("A")) if order == "jacobi" else (gr.operator("A"), gr.operator("C"))
    net = Net(sg.num_users, sg.num_items, d, K, *opsv).to(DEV)
    e0u, e0i = net.user_emb.weight.detach().cpu().numpy(), net.item_emb.weight.detach().cpu().numpy()
    users, pos, neg = _triples(sg, batch, seed=7)
    step = m.TrainStep(net, reg_weight=1e-4)
    loss = step.forward_backward(torch.tensor(users), torch.tensor(pos), torch.tensor(neg))
    o_loss, o_gu, o_gi, o_fu, o_fi = orc.train_step_grads(ops, e0u, e0i, users, pos, neg, K, order, 1e-4)
    errs = dict(loss=abs(float(loss.item()) - o_loss) / abs(o_loss),
                f_u=rel_err(step.f_u.cpu().numpy(), o_fu), f_i=rel_err(step.f_i.cpu().numpy(), o_fi),
                g_u=rel_err(net.user_emb.weight.grad.cpu().numpy(), o_gu),
                g_i=rel_err(net.item_emb.weight.grad.cpu().numpy(), o_gi))
    assert all(v <= TOL for v in errs.values()), errs
    # top-20 of a 1,000-user subsample on the propagated tables: fp32 kernel and tensor-core kernel vs the oracle
    rng = np.random.default_rng(3)
    ev_users = np.sort(rng.choice(sg.num_users, size=min(n_eval, sg.num_users), replace=False))
    tr = orc.edges_to_user_csr(sg.train_edges, sg.num_users)
    fu, fi = step.f_u, step.f_i
    o_ids, o_sc = orc.full_rank_topk(fu.cpu().numpy(), fi.cpu().numpy(), ev_users, tr, 20)
    for prec in ("fp32", "bf16x3"):
        ids, sc = ev.topk_device(fu, fi, torch.tensor(ev_users), (gr.samp_indptr, gr.samp_idx), 20, prec)
        _check_topk(ids.cpu().numpy(), sc.cpu().numpy(), o_ids, o_sc)
    return errs


def test_c2_full_size_against_the_oracle(cg):
    sg = cg["synth"].make_graph("C2")
    gr = cg["graph"].build_graph(sg.train_edges, sg.num_users, sg.num_items, sg.cred, "v2", DEV)
    ops = orc.Operators(sg.train_edges, sg.num_users, sg.num_items, sg.cred, "v2")
    _step_parity(cg, sg, gr, ops, 64, 3, "gs")


def test_c3_full_size_degree_aware_four_layers_against_the_oracle(cg):
    sg = cg["synth"].make_graph("C3")
    gr = cg["graph"].build_graph(sg.train_edges, sg.num_users, sg.num_items, sg.cred, "da", DEV)
    ops = orc.Operators(sg.train_edges, sg.num_users, sg.num_items, sg.cred, "da")
    _step_parity(cg, sg, gr, ops, 64, 4, "gs")


def test_c1_shape_jacobi_order_with_fairness_against_the_oracle(cg):
    """BASELINE config 1 at its own size through the same checks (CU operator, Jacobi order)."""
    sg = cg["synth"].make_graph("C1", duplicate_edges=200)
    gr = cg["graph"].build_graph(sg.train_edges, sg.num_users, sg.num_items, sg.cred, "cu", DEV)
    ops = orc.Operators(sg.train_edges, sg.num_users, sg.num_items, sg.cred, "cu")
    _step_parity(cg, sg, gr, ops, 64, 3, "jacobi", batch=943, n_eval=943)


def test_c4_shaped_subsample_crosses_the_regime_switch_against_the_oracle(cg):
    U, I, E, d, K = 625_000, 250_000, 12_500_000, 128, 3
    sg = cg["synth"].make_graph("C4", num_users=U, num_items=I, num_edges=E)
    gr = cg["graph"].build_graph(sg.train_edges, U, I, sg.cred, "v2", DEV)
    l2 = cg["lib"].get_option("L2_TABLE_BYTES")
    assert U * d * 4 > l2 and I * d * 4 > l2, "both gathered tables must exceed the regime switch on their own"
    assert gr.by_item.n_huge >= 1 and int(gr.deg_i.max()) > 16384, "needs a row for the finishing kernel"
    gr.set_emb_dim(d)
    assert gr.by_user.idx_hint is not None and gr.by_item.idx_hint is not None and gr.by_user.n_hot > 0
    hot = (gr.by_user.idx_hint < 0)
    assert torch.equal(gr.by_user.idx_hint & 0x7fffffff, gr.by_user.idx) and 0 < int(hot.sum()) < gr.nnz
    ops = orc.Operators(sg.train_edges, U, I, sg.cred, "v2")
    errs = _step_parity(cg, sg, gr, ops, d, K, "gs", n_eval=1000, check_graph=True)
    print("C4s rel errors", errs)


@pytest.mark.parametrize("d", [64, 128, 256])
def test_spmm_forms_give_identical_bits(cg, d):
    """The HBM-resident geometry (32 bytes per lane, 256-bit accesses) with and without hot-row hints and the
    L2-resident geometry all add the same terms in the same order: every combination must reproduce the same bits --
    plain and adjoint products, zero-degree rows, chunked and huge rows, both propagation orders."""
    lib, m = cg["lib"], cg["model"]
    sg = cg["synth"].make_graph("C1", num_users=6000, num_items=900, num_edges=400_000, duplicate_edges=500)
    gr = cg["graph"].build_graph(sg.train_edges, sg.num_users, sg.num_items, sg.cred, "v2", DEV)
    assert gr.by_item.n_huge >= 0 and gr.by_item.n_long > 0
    gen = torch.Generator(device=DEV).manual_seed(d)
    xu = torch.randn(sg.num_users, d, device=DEV, generator=gen)
    xi = torch.randn(sg.num_items, d, device=DEV, generator=gen)
    gs = torch.zeros_like(xu)
    gs[::37] = xu[::37]

    def run():
        return (m.spmm(gr.by_user, xi), m.spmm(gr.by_item, xu), m.spmm(gr.by_user, xi, use_bwd_values=True),
                m.spmm(gr.by_item, xu, use_bwd_values=True), *m.propagate_forward(gr, xu, xi, 3, "gs"),
                *m.propagate_forward(gr, xu, xi, 2, "jacobi"), *m.propagate_backward(gr, gs, xi, 3, "gs"),
                *m.propagate_backward(gr, gs, xi, 2, "jacobi"))

    want = run()                                  # L2-resident geometry, no hints
    keep = {k: lib.get_option(k) for k in ("L2_TABLE_BYTES", "HOT_ROWS")}
    try:
        lib.set_option("L2_TABLE_BYTES", 0)       # every table counts as HBM-resident
        for use_hints in (1, 0):
            for n_hot in (0, 64):
                lib.set_option("HOT_ROWS", use_hints)
                gr.by_user.set_hot_columns(gr.by_item.perm, n_hot)
                gr.by_item.set_hot_columns(gr.by_user.perm, 8 * n_hot)
                for a, b in zip(run(), want):
                    assert torch.equal(a.view(torch.int32), b.view(torch.int32)), (use_hints, n_hot)
    finally:
        for k, v in keep.items():
            lib.set_option(k, v)
        gr.by_user.set_hot_columns(gr.by_item.perm, 0)
        gr.by_item.set_hot_columns(gr.by_user.perm, 0)


def test_eval_scanning_groups_give_identical_results(cg):
    """CGX_OPT_EVAL_GROUPS: the tensor-core evaluation kernel with one scanning warp group (one list of 32 candidates
    per user) and with two (two lists of 24 over four TMEM accumulators, the default) must both return what the fp32
    kernel returns -- same ids, same score bits -- with exact score ties across many items and masked train items."""
    lib, ev = cg["lib"], cg["evaluate"]
    U, I, d = 700, 9000, 64
    sg = cg["synth"].make_graph("C1", num_users=U, num_items=I, num_edges=60_000)
    rng = np.random.default_rng(11)
    fu = torch.tensor((rng.standard_normal((U, d)) * 0.2).astype(np.float32), device=DEV)
    fi = torch.tensor((rng.standard_normal((I, d)) * 0.2).astype(np.float32), device=DEV)
    fi[::7] = fi[3]
    csr = ev._device_csr(orc.edges_to_user_csr(sg.train_edges, U), DEV)
    users = torch.arange(0, U, 2)
    ids0, sc0 = ev.topk_device(fu, fi, users, csr, 20, "fp32")
    keep = lib.get_option("EVAL_GROUPS")
    try:
        for groups in (1, 2):
            lib.set_option("EVAL_GROUPS", groups)
            ids1, sc1 = ev.topk_device(fu, fi, users, csr, 20, "bf16x3")
            assert torch.equal(ids0, ids1), (groups, int((ids0 != ids1).sum()))
            assert torch.equal(sc0, sc1), groups
    finally:
        lib.set_option("EVAL_GROUPS", keep)
