"""Generate golden fixtures by RUNNING THE REFERENCE (imported by path from /root/reference).

Run in the build container only (the GPU box has no /root/reference):
    python tests/golden/make_golden.py
Writes tests/golden/<case>_<variant>.npz.  Recorded with torch 2.11.0+cu128 / numpy 2.3.5.
"""
import importlib.util
import io
import contextlib
import pathlib
import sys

import numpy as np
import torch

ROOT = pathlib.Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from credgcn import synth  # noqa: E402

REF = pathlib.Path("/root/reference")
FILES = {
    "cu": REF / "lightgcn_cu.py",
    "v2": REF / "Version-2" / "lighgcn_cu_pop.py",
    "da": REF / "version_1" / "lightgcn_cu_pop_Degree-Aware Message.py",
}
CASES = {
    "tiny": dict(num_users=50, num_items=70, num_edges=420, duplicate_edges=12, emb_dim=64, batch=40),
    "small": dict(num_users=300, num_items=420, num_edges=6000, duplicate_edges=25, emb_dim=32, batch=256),
}
LAYERS = {"cu": 3, "v2": 3, "da": 4}


def load(tag):
    spec = importlib.util.spec_from_file_location(f"ref_{tag}", FILES[tag])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.cfg.device = "cpu"
    return mod


def coo(m):
    m = m.coalesce()
    return m.indices().numpy().copy(), m.values().numpy().copy()


def run(case, tag):
    c = CASES[case]
    g = synth.make_graph("C1", num_users=c["num_users"], num_items=c["num_items"], num_edges=c["num_edges"],
                         duplicate_edges=c["duplicate_edges"], seed=1234 + len(case))
    ref = load(tag)
    U, I, K, d = g.num_users, g.num_items, LAYERS[tag], c["emb_dim"]
    out = dict(train_edges=g.train_edges, val_edges=g.val_edges, test_edges=g.test_edges, cred=g.cred,
               num_users=U, num_items=I, num_layers=K, emb_dim=d)

    tr_csr = ref.edges_to_user_csr(g.train_edges, U)
    te_csr = ref.edges_to_user_csr(g.test_edges, U)
    out["csr_indptr"], out["csr_indices"] = tr_csr
    out["test_indptr"], out["test_indices"] = te_csr

    with contextlib.redirect_stdout(io.StringIO()):
        if tag == "cu":
            C, A, deg_i = ref.build_cred_weighted_mats(g.train_edges, U, I, g.cred, "cpu")
            out["deg_i"] = deg_i
        else:
            A, C = ref.build_message_passing_mats(g.train_edges, U, I, torch.tensor(g.cred), "cpu")
    out["A_idx"], out["A_val"] = coo(A)        # [U x I] base operator
    out["C_idx"], out["C_val"] = coo(C)        # [I x U] credibility operator

    torch.manual_seed(42)
    if tag == "cu":
        model = ref.CredLightGCN(U, I, d, K, C, A)
    else:
        model = ref.LightGCN(U, I, d, K, A, C)
    out["e0_u"] = model.user_emb.weight.detach().numpy().copy()
    out["e0_i"] = model.item_emb.weight.detach().numpy().copy()

    users, pos, neg = synth.make_triples(g, c["batch"], seed=5)
    out["users"], out["pos"], out["neg"] = users, pos, neg
    ut, pt, nt = (torch.tensor(x) for x in (users, pos, neg))

    def step(fair):
        model.zero_grad()
        if tag == "cu":
            eu, ei = model.final_embeddings()
            ps, ns = model.score(ut, pt, eu, ei), model.score(ut, nt, eu, ei)
            loss = -torch.log(torch.sigmoid(ps - ns) + 1e-12).mean()
            pop = torch.tensor((deg_i / max(float(deg_i.max()), 1.0)).astype(np.float32))
            loss = loss + fair * (pop[pt] * ps).mean() + ref.cfg.lambda_reg * model.l2_reg(ut, pt, nt)
        else:
            eu, ei = model.get_user_item_emb()
            loss = model.bpr_loss(ut, pt, nt, eu, ei, ref.cfg.reg)
        loss.backward()
        return (eu.detach().numpy().copy(), ei.detach().numpy().copy(), float(loss.item()),
                model.user_emb.weight.grad.numpy().copy(), model.item_emb.weight.grad.numpy().copy())

    out["final_u"], out["final_i"], out["loss"], out["grad_u"], out["grad_i"] = step(0.0)
    if tag == "cu":
        _, _, out["loss_fair"], out["grad_u_fair"], out["grad_i_fair"] = step(0.01)

    # samplers under a fixed seed (CU:615-622 / V2:835-849)
    rng = np.random.default_rng(ref.cfg.seed)
    indptr, indices = tr_csr
    train_users = np.where((indptr[1:] - indptr[:-1]) > 0)[0]
    rng.shuffle(train_users)
    batch = train_users[: c["batch"]]
    su, sp_, sn = [], [], []
    if hasattr(ref, "sample_neg_item_popmix"):
        item_deg = np.bincount(g.train_edges[1].astype(np.int64), minlength=I).astype(np.float64)
        popw = np.power(item_deg + 1.0, ref.cfg.neg_pop_gamma)
        pop_prob = (popw / (popw.sum() + 1e-12)).astype(np.float64)
        out["pop_prob"] = pop_prob
    for u in batch:
        p = ref.sample_pos_item(indptr, indices, int(u), rng)
        if p is None:
            continue
        if hasattr(ref, "sample_neg_item_popmix"):
            n = ref.sample_neg_item_popmix(indptr, indices, int(u), I, rng, pop_prob=pop_prob,
                                           mix_pop=ref.cfg.neg_mix_pop, max_tries=ref.cfg.neg_max_tries)
        else:
            n = ref.sample_neg_item(indptr, indices, int(u), I, rng)
        su.append(int(u)); sp_.append(p); sn.append(n)
    out["shuffled_users"] = train_users
    out["samp_users"], out["samp_pos"], out["samp_neg"] = map(np.asarray, (su, sp_, sn))

    # evaluation: record every ranked list the reference hands to metrics_at_k
    logged = []
    orig = ref.metrics_at_k

    def spy(ranked, gt, Kk):
        if Kk == max(ref.cfg.Ks):
            logged.append(np.asarray(ranked[:Kk]).astype(np.int64).copy())
        return orig(ranked, gt, Kk)

    ref.metrics_at_k = spy
    model.eval()
    if tag == "cu":
        res = ref.evaluate_sampled(model, tr_csr, te_csr, I, "cpu")
        out["sampled_ranked"] = np.stack(logged)
        keys = ("precision", "recall", "ndcg")
    else:
        item_pop, total = ref.compute_item_popularity(g.train_edges, I) if hasattr(ref, "compute_item_popularity") \
            else (None, 0)
        if tag == "v2":
            res = ref.evaluate_sampled(model, tr_csr, te_csr, I, "cpu", item_pop, total, g.cred)
            keys = ("precision", "recall", "ndcg", "item_coverage", "avg_log_popularity",
                    "avg_self_information", "cred_utility", "high_cred_recall", "low_cred_recall")
        else:
            res = ref.evaluate_sampled(model, tr_csr, te_csr, I, "cpu")
            keys = ("precision", "recall", "ndcg")
        out["sampled_ranked"] = np.stack(logged)
        for K in ref.cfg.Ks:
            out[f"sampled_{K}"] = np.array([res[K][k] for k in keys], np.float64)
        logged.clear()
        if tag == "v2":
            res = ref.evaluate_full_ranking(model, tr_csr, te_csr, I, "cpu", item_pop, total, g.cred)
            out["item_pop"], out["total_train"] = item_pop, total
        else:
            res = ref.evaluate_full_ranking(model, tr_csr, te_csr, I, "cpu")
        out["full_ranked"] = np.stack(logged)
        for K in ref.cfg.Ks:
            out[f"full_{K}"] = np.array([res[K][k] for k in keys], np.float64)
    if tag == "cu":
        for K in ref.cfg.Ks:
            out[f"sampled_{K}"] = np.array([res[K][k] for k in keys], np.float64)
    ref.metrics_at_k = orig

    path = ROOT / "tests" / "golden" / f"{case}_{tag}.npz"
    np.savez_compressed(path, **out)
    print(path.name, {k: getattr(v, "shape", v) for k, v in out.items() if k in ("A_val", "final_u", "loss")})


if __name__ == "__main__":
    for case in CASES:
        for tag in FILES:
            run(case, tag)
