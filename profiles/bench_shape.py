"""Propagation timing on an arbitrary power-law shape, per product kind:  python profiles/bench_shape.py U I E d [K]

Builds the synthetic graph on the device (the C4/C5 generator law), then times the item-row product (C x_u), the
user-row product (A x_i, with the running-sum epilogue) and a whole forward + adjoint propagation with CUDA events,
with / without the hot-row hints.  Used for the C5-shard shape (6.25M x 10M x 125M, d = 64): the per-rank products of the 8-GPU run."""
import json
import pathlib
import sys

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import torch
from credgcn import _lib, graph, model, synth

U, I, E, d = (int(x) for x in sys.argv[1:5])
K = int(sys.argv[5]) if len(sys.argv) > 5 else 3
dev = torch.device("cuda", 0)
sg = synth.make_graph_device("C5", dev, num_users=U, num_items=I, num_edges=E)
gr = graph.build_graph(sg.train_edges, U, I, sg.cred, "v2", dev)
nnz = gr.nnz
del sg
torch.cuda.empty_cache()
gen = torch.Generator(device=dev).manual_seed(0)
xu = torch.randn(U, d, device=dev, generator=gen) * 0.1
xi = torch.randn(I, d, device=dev, generator=gen) * 0.1
acc = torch.zeros_like(xu)
out = {"shape": [U, I, nnz, d], "rows": {"user_avg_nnz": nnz / U, "item_avg_nnz": nnz / I}}


def timed(fn, n=5):
    for _ in range(2):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


for hot in (1, 0):
    gr.set_emb_dim(d, hot_bytes=None if hot else 0)
    for _ in (0,):
        key = "hot" if hot else "nohot"
        out[key] = {
            "item_rows_ms": timed(lambda: model.spmm(gr.by_item, xu)),
            "user_rows_ms": timed(lambda: model.spmm(gr.by_user, xi)),
            "fwd_ms": timed(lambda: model.propagate_forward(gr, xu, xi, K, "gs"), 3),
            "bwd_ms": timed(lambda: model.propagate_backward(gr, acc, xi, K, "gs"), 3),
        }
        r = 4 * d
        out[key]["item_rows_gather_model_gbs"] = (nnz * (8 + r) + I * r) / out[key]["item_rows_ms"] / 1e6
        out[key]["user_rows_gather_model_gbs"] = (nnz * (8 + r) + U * r) / out[key]["user_rows_ms"] / 1e6
print(json.dumps(out))
