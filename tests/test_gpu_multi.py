"""Multi-GPU checks of the user-sharded path: sharded forward / loss / backward on 2, 4 and 8 ranks == the
single-GPU path on the whole graph (<= 1e-4 relative), for all three operator variants (Gauss-Seidel and Jacobi
order), with the NCCL exchange, the peer-memory pull kernel and the pushed form (rows pushed from the SpMM epilogue,
the default above 2 ranks); user-sharded full-rank evaluation == single-GPU evaluation.  Each case is skipped when
the box has fewer GPUs (`gpurun --gpus 8 -- python -m pytest tests/test_gpu_multi.py -m gpu`)."""
import os
import pathlib
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = pathlib.Path(__file__).resolve().parents[1]
for _p in (ROOT, ROOT / "oracle"):
    if str(_p) not in sys.path:
        sys.path.insert(0, str(_p))
pytestmark = pytest.mark.gpu


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from credgcn import synth
        from credgcn.sharded import (CollectiveExchange, P2PExchange, ShardedTrainStep, build_local_graph,
                                     parity_vs_single_gpu, partition_users, shard_edges)
        I, d = synth.SHAPES["C1"]["num_items"], 64
        res = {}
        p2p = P2PExchange(2 * I * d + 4, dev, backing="auto",   # NVLink peer-memory exchange (csrc/comm.cu)
                          gather_floats=ShardedTrainStep._block_floats(2048, d))
        for variant, order in (("v2", "gs"), ("cu", "jacobi"), ("da", "gs")):
            outs = {}
            cases = [("nccl", CollectiveExchange(), None, False), ("p2p", p2p, False, False),
                     ("p2p_push", p2p, True, False)]
            if p2p.mc:                        # NVSwitch multicast mapping available: the NVLS form of the exchange
                cases.append(("p2p_nvls", p2p, False, True))
            for ex_name, ex, push, nvls in cases:
                p2p.force_push, p2p.use_nvls = push, nvls
                worst, errs, tensors = parity_vs_single_gpu(rank, world, dev, variant, order, exchange=ex)
                outs[ex_name] = [t.clone() for t in tensors]
                res[f"{variant}/{ex_name}"] = worst
            p2p.force_push, p2p.use_nvls = None, None
            # the pull kernel and the pushed form add the partials in rank order: identical bits; NCCL's order (and
            # the switch's, in the NVLS form) is only the same at two ranks
            same = all(torch.equal(a, b) for a, b in zip(outs["p2p"], outs["p2p_push"]))
            if world == 2:
                same = same and all(torch.equal(a, b) for a, b in zip(outs["nccl"], outs["p2p"]))
                if "p2p_nvls" in outs:
                    same = same and all(torch.equal(a, b) for a, b in zip(outs["p2p_nvls"], outs["p2p"]))
            flag = torch.tensor([1.0 if same else 0.0], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            res[f"{variant}/bitwise"] = 0.0 if flag.item() == 1.0 else 1.0
        res["nvls_cases_ran"] = 0.0
        if rank == 0:
            print(f"[multi] world={world} backing={'symmetric memory' if p2p._symm is not None else 'cudaIpc'} "
                  f"multicast={'yes' if p2p.mc else 'no'}", flush=True)
        p2p.check()

        # user-sharded full-rank evaluation == single-GPU evaluation of the whole graph
        from credgcn import evaluate
        from credgcn.sharded import evaluate_full_ranking_sharded
        sg = synth.make_graph("C1", duplicate_edges=100)
        U = sg.num_users
        bounds = partition_users(np.bincount(sg.train_edges[0], minlength=U), world)
        lo, hi = int(bounds[rank]), int(bounds[rank + 1])
        gl = build_local_graph(shard_edges(sg.train_edges, bounds, rank), hi - lo, I, sg.cred[lo:hi], "v2", dev)
        torch.manual_seed(3)
        fu = torch.randn(U, d, device=dev) * 0.2
        fi = torch.randn(I, d, device=dev) * 0.2
        item_pop, total = evaluate.compute_item_popularity(sg.train_edges, I)
        got = evaluate_full_ranking_sharded(fu[lo:hi].contiguous(), fi, gl, shard_edges(sg.test_edges, bounds, rank), I,
                                            (10, 20), item_pop, total, sg.cred[lo:hi])
        if rank == 0:
            class _M:                      # minimal stand-in with the reference's accessor
                def get_user_item_emb(self):
                    return fu, fi
            import credgcn_oracle as orc
            want = evaluate.evaluate_full_ranking(_M(), orc.edges_to_user_csr(sg.train_edges, U),
                                                  orc.edges_to_user_csr(sg.test_edges, U), I, dev, item_pop, total, sg.cred)
            ev_err = 0.0
            for kk in (10, 20):
                for key in ("precision", "recall", "ndcg", "item_coverage", "avg_log_popularity",
                            "avg_self_information", "cred_utility", "high_cred_recall", "low_cred_recall"):
                    ev_err = max(ev_err, abs(got[kk][key] - want[kk][key]) / max(abs(want[kk][key]), 1e-12))
                assert got[kk]["users_eval"] == want[kk]["users_eval"]
                assert (got[kk]["high_users"], got[kk]["low_users"]) == (want[kk]["high_users"], want[kk]["low_users"])
            res["eval"] = ev_err
            out[0] = res
        dist.barrier()
        p2p.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(900)
@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_equals_single_gpu(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    port = 29600 + (os.getpid() % 2000) + world
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        res = dict(out)[0]
    print(f"world={world}: " + ", ".join(f"{k}={v:.2e}" for k, v in sorted(res.items())))
    for key, err in res.items():
        assert err < 1e-4, (world, key, err)
