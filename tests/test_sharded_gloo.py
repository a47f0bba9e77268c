"""Host-side logic of the user-sharded path on CPU: world_size-2 and -4 gloo process groups, with an
oracle-backed stand-in for the two per-shard products (the product backend is CUDA only).
Checks that the sharded forward / adjoint schedules with their all-reduces reproduce the
single-process oracle, for both layer orders."""
import os
import sys
import pathlib

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = pathlib.Path(__file__).resolve().parents[1]
for p in (ROOT, ROOT / "oracle"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))


def test_partition_users_balances_nonzeros():
    from credgcn.sharded import partition_users, shard_edges
    rng = np.random.default_rng(0)
    deg = rng.zipf(1.6, size=5000).clip(max=3000)
    for world in (1, 2, 3, 8):
        b = partition_users(deg, world)
        assert b[0] == 0 and b[-1] == deg.size and (np.diff(b) >= 0).all() and len(b) == world + 1
        loads = np.array([deg[b[r]:b[r + 1]].sum() for r in range(world)])
        assert loads.sum() == deg.sum()
        assert loads.max() <= deg.sum() / world + deg.max()          # within one row of the ideal split
    edges = np.stack([rng.integers(0, 100, 1000), rng.integers(0, 50, 1000)]).astype(np.int32)
    b = partition_users(np.bincount(edges[0], minlength=100), 4)
    parts = [shard_edges(edges, b, r) for r in range(4)]
    assert sum(p.shape[1] for p in parts) == 1000
    for r, p in enumerate(parts):
        assert p.shape[1] == 0 or (p[0].min() >= 0 and p[0].max() < b[r + 1] - b[r])


from credgcn.sharded import BackendBase  # noqa: E402


class OracleShardBackend(BackendBase):
    """Test double of sharded.CudaBackend: SciPy products of one user shard, weights from GLOBAL degrees."""

    def __init__(self, edges, lo, hi, num_items, cred, variant, deg_i_global):
        import credgcn_oracle as orc
        u = edges[0].astype(np.int64)
        keep = (u >= lo) & (u < hi)
        ul, il = u[keep] - lo, edges[1].astype(np.int64)[keep]
        deg_u = np.bincount(ul, minlength=hi - lo).astype(np.float32)
        w_a, w_c = orc.edge_weights(variant, ul, il, deg_u, deg_i_global, cred[lo:hi])
        import scipy.sparse as sp
        ar, ac, av = orc.coalesce(ul, il, w_a)
        cr, cc, cv = orc.coalesce(il, ul, w_c)
        self.A = sp.csr_matrix((av, (ar, ac)), shape=(hi - lo, num_items), dtype=np.float32)
        self.C = sp.csr_matrix((cv, (cr, cc)), shape=(num_items, hi - lo), dtype=np.float32)

    def item_rows(self, x_u, bwd=False, out=None):
        m = self.A.T if bwd else self.C
        res = torch.from_numpy(np.asarray(m @ x_u.numpy(), dtype=np.float32))
        return res if out is None else out.copy_(res)

    def user_rows(self, x_i, bwd=False):
        m = self.C.T if bwd else self.A
        return torch.from_numpy(np.asarray(m @ x_i.numpy(), dtype=np.float32))


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import credgcn_oracle as orc
        from credgcn import synth
        from credgcn.sharded import ShardedPropagation, partition_users
        sg = synth.make_graph("C1", num_users=240, num_items=150, num_edges=6000, duplicate_edges=30)
        U, I, d, K = sg.num_users, sg.num_items, 16, 3
        deg_u, deg_i = orc.degrees(sg.train_edges, U, I)
        bounds = partition_users(deg_u, world)
        lo, hi = int(bounds[rank]), int(bounds[rank + 1])
        rng = np.random.default_rng(5)
        eu = (rng.standard_normal((U, d)) * 0.1).astype(np.float32)
        ei = (rng.standard_normal((I, d)) * 0.1).astype(np.float32)
        gu = rng.standard_normal((U, d)).astype(np.float32)
        gi = rng.standard_normal((I, d)).astype(np.float32)
        errs = {}
        for variant, order in (("cu", "jacobi"), ("v2", "gs"), ("da", "gs")):
            ops = orc.Operators(sg.train_edges, U, I, sg.cred, variant)
            be = OracleShardBackend(sg.train_edges, lo, hi, I, sg.cred, variant, deg_i)
            prop = ShardedPropagation(be, K, order)
            fu, fi = prop.forward(torch.from_numpy(eu[lo:hi].copy()), torch.from_numpy(ei.copy()))
            wu, wi = orc.propagate(ops, eu, ei, K, order)
            # each rank contributes gi / world to the item seed; the caller reduces it
            g_i_total = torch.from_numpy(gi / world)
            dist.all_reduce(g_i_total)
            bu, bi = prop.backward(torch.from_numpy(gu[lo:hi].copy()), g_i_total)
            ou, oi = orc.propagate_backward(ops, gu, gi, K, order)
            rel = lambda a, b: float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))
            errs[variant] = (rel(fu.numpy(), wu[lo:hi]), rel(fi.numpy(), wi), rel(bu.numpy(), ou[lo:hi]),
                             rel(bi.numpy(), oi))
        out[rank] = errs
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(240)
@pytest.mark.parametrize("world", [2, 4])
def test_sharded_propagation_matches_single_process(world):
    port = 29500 + (os.getpid() % 2000) + 11 * world
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        res = dict(out)
    assert set(res) == set(range(world))
    for rank, errs in res.items():
        for variant, e in errs.items():
            assert max(e) < 1e-5, (rank, variant, e)


def _gather_worker(rank, world, port, out):
    """Compact loss-gradient exchange of ShardedTrainStep on the CPU: pack -> all-gather (gloo) -> unpack -> rank-order
    accumulation == an all-reduce of the dense tables; sparse seed addition of the adjoint == the dense addition."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from credgcn.sharded import (CollectiveExchange, ShardedPropagation, distinct_rows, item_gradient_block_floats,
                                     pack_item_gradient, unpack_item_gradient)
        U, I, d, Bm = 40, 60, 8, 16
        B = 16 if rank == 0 else 11                               # ranks may hold different batch sizes
        rng = np.random.default_rng(100 + rank)
        # what cgx_bpr_fwd_bwd leaves behind: a plan sorted by row (B user entries, then 2B item entries); ego_rows names
        # the row at the head of every run and is -1 elsewhere; ego_coef is only written at heads (NaN elsewhere here)
        items = np.sort(rng.integers(0, I, size=2 * B))
        head = np.concatenate([[True], items[1:] != items[:-1]])
        ego_rows = np.full(3 * B, -1, np.int32)
        ego_rows[:B] = np.arange(B)                               # user part (ignored by the packer)
        ego_rows[B:][head] = (U + items[head]).astype(np.int32)
        ego_coef = np.full(3 * B, np.nan, np.float32)
        ego_coef[B:][head] = rng.random(int(head.sum())).astype(np.float32)
        gi_local = np.zeros((I, d), np.float32)
        gi_local[items[head]] = rng.standard_normal((int(head.sum()), d)).astype(np.float32)
        loss = torch.tensor([0.25 * (rank + 1)])
        block = pack_item_gradient(torch.from_numpy(ego_rows), torch.from_numpy(ego_coef), torch.from_numpy(gi_local), loss,
                                   B, Bm, U, d)
        assert block.numel() == item_gradient_block_floats(Bm, d) and torch.isfinite(block[: 2 * Bm * (d + 1)]).all()
        blocks = CollectiveExchange().allgather(block)
        vals, coef, rows_all, ok, total_loss = unpack_item_gradient(blocks, Bm, d, I)
        assert abs(float(total_loss) - 0.25 * sum(range(1, world + 1))) < 1e-6
        assert int(ok[rank].sum()) == int(head.sum())
        seed = torch.zeros(I, d)
        l2 = torch.zeros(I, d)
        e0_i = torch.from_numpy(np.random.default_rng(7).standard_normal((I, d)).astype(np.float32))
        for r in range(world):
            rr = rows_all[r].clamp(max=I - 1)
            seed.index_add_(0, rr, torch.where(ok[r][:, None], vals[r], 0.0))
            l2.index_add_(0, rr, e0_i[rr] * (coef[r] * ok[r])[:, None])
        # dense truth: all-reduce of the per-rank dense tables
        want_seed = torch.from_numpy(gi_local.copy())
        c_dense = np.zeros(I, np.float32)
        c_dense[items[head]] = ego_coef[B:][head]
        want_l2 = e0_i * torch.from_numpy(c_dense)[:, None]
        dist.all_reduce(want_seed)
        dist.all_reduce(want_l2)
        err = max(float((seed - want_seed).abs().max()), float((l2 - want_l2).abs().max()))
        # one representative per distinct row; sparse seed addition == dense addition
        owner = torch.full((I + 1,), -1, dtype=torch.int64)
        flat_c, keep = distinct_rows(rows_all, ok, owner, I)
        touched = torch.unique(rows_all[ok])
        assert int(keep.sum()) == touched.numel() and set(flat_c[keep].tolist()) == set(touched.tolist())
        t = torch.from_numpy(np.random.default_rng(9).standard_normal((I, d)).astype(np.float32))
        sparse = t.clone().index_add_(0, flat_c, seed[flat_c] * keep[:, None])
        err = max(err, float((sparse - (t + seed)).abs().max()))
        out[rank] = err
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(240)
@pytest.mark.parametrize("world", [2, 4])
def test_compact_loss_gradient_exchange(world):
    port = 29500 + (os.getpid() % 2000) + 7 + 11 * world
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_gather_worker, args=(world, port, out), nprocs=world, join=True)
        res = dict(out)
    assert set(res) == set(range(world)) and max(res.values()) < 1e-6, res
