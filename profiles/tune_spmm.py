"""Time one K=3 forward + backward propagation for the SpMM variant in $CGX_SPMM_VARIANT.
    make -C <package>/csrc clean all EXTRA=-DCGX_SPMM_TUNING     (the variants are not in the default build)
    CGX_SPMM_VARIANT=n python profiles/tune_spmm.py <shape> <d>
shape: C2 | mid (4M x 1M x 64M, device generated) | C4"""
import json, os, pathlib, sys, time
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import torch
from credgcn import graph, model, synth

shape, d = sys.argv[1], int(sys.argv[2])
dev = torch.device("cuda", 0)
if shape == "C2":
    sg = synth.make_graph("C2")
elif shape == "mid":
    sg = synth.make_graph_device("C4", dev, num_users=4_000_000, num_items=1_000_000, num_edges=80_000_000)
else:
    sg = synth.make_graph_device("C4", dev)
gr = graph.build_graph(sg.train_edges, sg.num_users, sg.num_items, sg.cred, "v2", dev)
U, I, nnz, K = sg.num_users, sg.num_items, gr.nnz, 3
xu = torch.randn(U, d, device=dev) * 0.1
xi = torch.randn(I, d, device=dev) * 0.1
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(3):
    model.propagate_forward(gr, xu, xi, K, "gs"); model.propagate_backward(gr, xu, xi, K, "gs")
torch.cuda.synchronize()
ts = []
for rep in range(8):
    flush.fill_(rep)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    model.propagate_forward(gr, xu, xi, K, "gs"); model.propagate_backward(gr, xu, xi, K, "gs")
    b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
ms = sorted(ts)[len(ts) // 2]
r = 4 * d
alg = 2 * (K * (2 * nnz * (8 + r) + (U + I) * r) + 2 * K * (U + I) * r)
print(json.dumps({"variant": int(os.environ.get("CGX_SPMM_VARIANT", "0")), "shape": shape, "d": d, "nnz": nnz,
                  "fwd_bwd_ms": round(ms, 4), "gather_model_GBs": round(alg / ms / 1e6, 1),
                  "edges_per_s": round(nnz / ms * 1e3), "n_long": [gr.by_user.n_long, gr.by_item.n_long],
                  "n_huge": [gr.by_user.n_huge, gr.by_item.n_huge]}))
