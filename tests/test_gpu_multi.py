"""2-GPU check of the user-sharded path over NCCL: sharded forward / loss / backward on two
ranks == the single-GPU path on the whole graph (<= 1e-4 relative).  Skipped with < 2 GPUs
(`gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`)."""
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
        from credgcn import graph, model, synth
        from credgcn.sharded import (CollectiveExchange, CudaBackend, P2PExchange, ShardedPropagation,
                                     all_reduce_sum, build_local_graph, partition_users, shard_edges)
        sg = synth.make_graph("C1", duplicate_edges=100)
        U, I, d, K = sg.num_users, sg.num_items, 64, 3
        deg_u = np.bincount(sg.train_edges[0], minlength=U)
        bounds = partition_users(deg_u, world)
        lo, hi = int(bounds[rank]), int(bounds[rank + 1])
        users, pos, neg = synth.make_triples(sg, 2048)
        mine = (users >= lo) & (users < hi)
        torch.manual_seed(0)
        eu = torch.nn.init.xavier_uniform_(torch.empty(U, d))
        ei = torch.nn.init.xavier_uniform_(torch.empty(I, d))
        res = {}
        p2p = P2PExchange(2 * I * d + 4, dev)         # NVLink peer-memory exchange (csrc/comm.cu)
        for variant, order in (("v2", "gs"), ("cu", "jacobi"), ("da", "gs")):
          gl = build_local_graph(shard_edges(sg.train_edges, bounds, rank), hi - lo, I, sg.cred[lo:hi], variant, dev)
          outs = {}
          # p2p = pull kernel (partials read over NVLink); p2p_push = rows pushed to their owner from the SpMM epilogue
          for ex_name, ex in (("nccl", CollectiveExchange()), ("p2p", p2p), ("p2p_push", p2p)):
            os.environ["CGX_P2P_PUSH"] = "1" if ex_name == "p2p_push" else "0"
            prop = ShardedPropagation(CudaBackend(gl), K, order, exchange=ex)
            eu_l, ei_d = eu[lo:hi].to(dev).contiguous(), ei.to(dev)
            f_u, f_i = prop.forward(eu_l, ei_d)
            g_u = torch.zeros_like(eu_l)
            gi2 = ex.partial_buffer((2, I, d), dev).zero_()
            ego_u = torch.zeros_like(eu_l)
            loss, _, _, ego_rows, ego_coef = model.bpr_fused(gl, f_u, f_i, eu_l, ei_d, users[mine] - lo, pos[mine],
                                                             neg[mine], 1e-4, 0.0, None, g_u, gi2[0],
                                                             batch_total=len(users))
            model.apply_ego(gl, ego_rows, ego_coef, eu_l, ei_d, ego_u, gi2[1])
            gi2 = ex.reduce(gi2).clone()
            all_reduce_sum(loss)
            d_u, d_i = prop.backward(g_u, gi2[0])
            d_u, d_i = d_u + ego_u, d_i + gi2[1]
            outs[ex_name] = (f_u.clone(), f_i.clone(), d_u.clone(), d_i.clone())
          os.environ.pop("CGX_P2P_PUSH", None)
          same = (all(torch.equal(a, b) for a, b in zip(outs["nccl"], outs["p2p"])) and
                  all(torch.equal(a, b) for a, b in zip(outs["nccl"], outs["p2p_push"])))
          if True:
            if rank == 0:      # single-GPU truth on the whole graph
                gr = graph.build_graph(sg.train_edges, U, I, sg.cred, variant, dev)
                Net = model.CredLightGCN if variant == "cu" else model.LightGCN
                ops = (gr.operator("C"), gr.operator("A")) if variant == "cu" else (gr.operator("A"), gr.operator("C"))
                net = Net(U, I, d, K, *ops)
                net.load_state_dict({"user_emb.weight": eu, "item_emb.weight": ei})
                net = net.to(dev)
                st = model.TrainStep(net, reg_weight=1e-4)
                want = st.forward_backward(users, pos, neg)
                rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
                res[variant] = dict(
                    loss=abs(loss.item() - want.item()) / abs(want.item()),
                    f_u=rel(f_u, st.f_u[lo:hi]), f_i=rel(f_i, st.f_i),
                    d_u=rel(d_u, net.user_emb.weight.grad[lo:hi]), d_i=rel(d_i, net.item_emb.weight.grad),
                    deg=int((gl.deg_i != gr.deg_i).sum().item()), p2p_equals_nccl=bool(same))
        # user-sharded full-rank evaluation == single-GPU evaluation of the whole graph
        from credgcn import evaluate, config
        from credgcn.sharded import evaluate_full_ranking_sharded
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
            res["eval"] = dict(err=ev_err, deg=0, p2p_equals_nccl=True)
            out[0] = res
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_gpu_sharded_equals_single_gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    port = 29600 + (os.getpid() % 2000)
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
        res = dict(out)[0]
    for variant, e in res.items():
        assert e.pop("deg") == 0, variant
        assert e.pop("p2p_equals_nccl"), variant          # two ranks: a + b in rank order == NCCL's sum, bit for bit
        assert max(e.values()) < 1e-4, (variant, e)
