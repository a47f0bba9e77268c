"""Full-rank top-K evaluation timing: fp32 CUDA-core path vs tcgen05 paths.  python profiles/bench_eval.py [C2|C3] [K]"""
import json, os, pathlib, sys
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import numpy as np, torch
from credgcn import evaluate, graph, model, synth

name = sys.argv[1] if len(sys.argv) > 1 else "C2"
dev = torch.device("cuda", 0)
torch.manual_seed(0)
if name.isdigit():      # python profiles/bench_eval.py U I d [K]: random tables, every user owns 20 random train items
    U, I, d = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    K = int(sys.argv[4]) if len(sys.argv) > 4 else 20
    fu = torch.randn(U, d, device=dev) * 0.1
    fi = torch.randn(I, d, device=dev) * 0.1
    idx = torch.sort(torch.randint(0, I, (U, 20), device=dev, dtype=torch.int32), dim=1).values.reshape(-1)
    csr = (torch.arange(0, 20 * U + 1, 20, device=dev, dtype=torch.int64), idx.contiguous())
    name = f"random {U}x{I}"

    class sg:
        num_users, num_items = U, I
else:
    K = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    sg = synth.make_graph(name)
    shp = synth.SHAPES[name]
    gr = graph.build_graph(sg.train_edges, sg.num_users, sg.num_items, sg.cred, shp["variant"], dev)
    d = shp["emb_dim"]
    eu = torch.nn.init.xavier_uniform_(torch.empty(sg.num_users, d)).to(dev)
    ei = torch.nn.init.xavier_uniform_(torch.empty(sg.num_items, d)).to(dev)
    fu, fi = model.propagate_forward(gr, eu, ei, shp["num_layers"], shp["order"])
    csr = (gr.samp_indptr, gr.samp_idx)
users = torch.arange(sg.num_users, device=dev)
out = {"workload": name, "users": sg.num_users, "items": sg.num_items, "d": d, "K": K,
       "flops": 2.0 * sg.num_users * sg.num_items * d}
ref = None
for prec in ("fp32", "bf16x3", "bf16"):
    for _ in range(2):
        ids, sc = evaluate.topk_device(fu, fi, users, csr, K, prec)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); ids, sc = evaluate.topk_device(fu, fi, users, csr, K, prec); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = sorted(ts)[2]
    mult = 3 if prec == "bf16x3" else 1
    out[prec] = {"ms": round(ms, 3), "useful_TFLOPs": round(out["flops"] / ms / 1e9, 1),
                 "tensor_TFLOPs_issued": round(mult * out["flops"] / ms / 1e9, 1) if prec != "fp32" else None}
    if ref is None:
        ref = ids
    else:
        out[prec]["ids_equal_fp32"] = bool(torch.equal(ref, ids))
        out[prec]["id_agreement"] = float((ref == ids).float().mean())
print(json.dumps(out))
