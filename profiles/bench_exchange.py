"""Item-table exchange alone: P2P kernel vs NCCL all-reduce, 2+ GPUs (torchrun)."""
import os, sys, pathlib, json
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
from credgcn.sharded import P2PExchange, CollectiveExchange
out = {}
for n in (38048 * 64, 2_000_000 * 128, 10_000_000 * 64):
    p2p = P2PExchange(n, dev, backing="auto")
    cases = [("p2p_pull", p2p, False), ("nccl", CollectiveExchange(), None)] + ([("p2p_nvls", p2p, True)] if p2p.mc else [])
    for name, ex, nvls in cases:
        p2p.use_nvls = nvls
        for _ in range(5):
            b = ex.partial_buffer((n,), dev); b.fill_(1.0); r = ex.reduce(b)
        torch.cuda.synchronize(); dist.barrier()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20):
            b = ex.partial_buffer((n,), dev); r = ex.reduce(b)
        e.record(); torch.cuda.synchronize()
        out[f"{name}_{n*4/1e6:.1f}MB_us"] = round(a.elapsed_time(e) * 50, 1)
        if name == "p2p_pull" and os.environ.get("CGX_OPT_P2P_TIMING"):
            import ctypes
            from credgcn._lib import lib
            t = (ctypes.c_uint64 * 4)()
            lib().cgx_comm_timing(t)
            if t[3]:
                out[f"p2p_{n*4/1e6:.1f}MB_phases_us"] = {"barrier_a": round(t[0] / t[3] / 1e3, 1),
                                                        "reduce": round(t[1] / t[3] / 1e3, 1),
                                                        "barrier_b": round(t[2] / t[3] / 1e3, 1), "n": int(t[3])}
    out[f"backing_{n*4/1e6:.1f}MB"] = ("symmetric memory" if p2p._symm is not None else "cudaIpc") + (", multicast" if p2p.mc else "")
    dist.barrier(); p2p.close(); del p2p
if rank == 0: print(json.dumps({"world": world, **out}))
dist.barrier(); dist.destroy_process_group()
