"""The reference arm of bench.py: oracle/ref_runner.py drives the UNMODIFIED reference scripts staged under oracle/_ref/
(oracle/make_ref.py).  These CPU tests check that the staged files are byte-identical to their manifest, that one
training step of the reference's own classes gives the loss the oracle computes from the same weights and triples, and
that its Python sampler returns valid triples.  Skipped when the scripts are not staged (a checkout without
/root/reference)."""
import hashlib
import json
import pathlib
import sys

import numpy as np
import pytest
import torch

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "oracle"))
import credgcn_oracle as orc  # noqa: E402
import ref_runner  # noqa: E402

pytestmark = pytest.mark.skipif(not ref_runner.available(), reason="oracle/_ref is not staged (run oracle/make_ref.py)")


def test_staged_scripts_match_their_manifest():
    man = json.loads((ref_runner.REF_DIR / "MANIFEST.json").read_text())
    for name in ref_runner.MODULES.values():
        assert hashlib.md5((ref_runner.REF_DIR / name).read_bytes()).hexdigest() == man[name]["md5"], name
    assert not any(p.suffix == ".py" and p.name not in ref_runner.MODULES.values() for p in ref_runner.REF_DIR.iterdir())


@pytest.mark.parametrize("variant,order", [("v2", "gs"), ("cu", "jacobi"), ("da", "gs")])
def test_reference_step_agrees_with_the_oracle(variant, order):
    from credgcn import synth
    sg = synth.make_graph("C1", num_users=120, num_items=90, num_edges=1500, duplicate_edges=10, seed=5)
    d, K = 16, 3 if variant != "da" else 4
    arm = ref_runner.ReferenceArm(variant, sg.train_edges, sg.num_users, sg.num_items, sg.cred, d, K)
    e0u = arm.model.user_emb.weight.detach().numpy().copy()
    e0i = arm.model.item_emb.weight.detach().numpy().copy()
    users, pos, neg = synth.make_triples(sg, 64)
    reg = arm.cfg.lambda_reg if variant == "cu" else arm.cfg.reg
    loss = arm.step(torch.tensor(users), torch.tensor(pos), torch.tensor(neg))
    ops = orc.Operators(sg.train_edges, sg.num_users, sg.num_items, sg.cred, variant)
    o_loss, o_gu, o_gi, _, _ = orc.train_step_grads(ops, e0u, e0i, users, pos, neg, K, order, reg)
    assert abs(loss - o_loss) <= 1e-5 * abs(o_loss), (loss, o_loss)
    # Adam's first step moves every parameter with a non-zero gradient by lr against the gradient's sign
    du = arm.model.user_emb.weight.detach().numpy() - e0u
    moved = np.abs(o_gu) > 1e-5          # (|g| >> eps = 1e-8, so the step is lr to a part in 1e3)
    assert moved.any() and np.all(np.sign(du[moved]) == -np.sign(o_gu[moved]))
    assert np.allclose(np.abs(du[moved]), arm.cfg.lr, rtol=2e-2)


def test_reference_sampler_loop_returns_valid_triples():
    from credgcn import synth
    sg = synth.make_graph("C1", num_users=200, num_items=150, num_edges=3000, seed=6)
    arm = ref_runner.ReferenceArm("v2", sg.train_edges, sg.num_users, sg.num_items, sg.cred, 16, 2)
    batch = arm.batches(64)[0]
    u, p, n = (t.numpy() for t in arm.sample_batch(batch))
    indptr, indices = arm.train_csr
    assert len(u) == len(batch) and set(u) <= set(arm.train_users)
    for uu, pp, nn in zip(u, p, n):
        row = indices[indptr[uu]:indptr[uu + 1]]
        assert pp in row and nn not in row and 0 <= nn < sg.num_items
