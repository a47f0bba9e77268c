"""Cost of the sharded code path without any peer: ShardedTrainStep on ONE rank (world_size 1) vs TrainStep."""
import os, sys, pathlib, time
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import numpy as np, torch, torch.distributed as dist
os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29533")
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
dist.init_process_group("nccl", rank=0, world_size=1, device_id=dev)
from credgcn import synth, sharded
sg = synth.make_graph("C2"); shp = synth.SHAPES["C2"]
gr = sharded.build_local_graph(sg.train_edges, sg.num_users, sg.num_items, sg.cred, "v2", dev)
torch.manual_seed(0)
eu = torch.nn.init.xavier_uniform_(torch.empty(sg.num_users, 64)).to(dev)
ei = torch.nn.init.xavier_uniform_(torch.empty(sg.num_items, 64)).to(dev)
st = sharded.ShardedTrainStep(gr, eu, ei, 3, "gs")
users = torch.nonzero(gr.deg_u > 0).reshape(-1)[:4096].contiguous()
for _ in range(10): st.step(users)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); a.record()
for _ in range(100): st.step(users)
b.record(); torch.cuda.synchronize()
print("sharded path, world=1: %.3f ms/step (events)  %.3f ms/step (wall)" % (a.elapsed_time(b) / 100, (time.perf_counter() - t0) * 10))
dist.destroy_process_group()
