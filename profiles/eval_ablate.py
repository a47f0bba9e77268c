"""Where the tensor-core evaluation spends its time:  python profiles/eval_ablate.py C2 [C3 ...] [--once | --variants]

Times cgx_eval_topk (bf16x3) on the propagated tables of a BASELINE shape.

default: parts of k_eval_umma switched off through CGX_OPT_EVAL_DEBUG (diagnostics only; results are wrong with any
    of these bits set):
        option 0   the product
        option 2   no score ever enters a candidate list (scan + pipeline only)
        option 6   the same without the MMAs
        option 14  the same without the TMA loads (barrier skeleton + TMEM reads + scan)
--variants: one or two scanning warp groups (CGX_OPT_EVAL_GROUPS), the second checked against the ids and score bits
    of the first, one JSON line per variant as soon as it is measured.  (profiles/r2_eval_variants.jsonl also holds the
    two forms that left the code after that run: `lane == 0` instead of elect.sync around the MMA issue -- slower -- and
    an out-of-line merge -- no difference.)
--once: two calls of the product and nothing else (the command an `ncu -k regex:k_eval_umma --launch-skip 1 -c 1`
    capture wraps)."""
import json
import pathlib
import sys

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import torch
from credgcn import _lib, evaluate, graph, model, synth

once = "--once" in sys.argv
variants = "--variants" in sys.argv
names = [a for a in sys.argv[1:] if not a.startswith("--")] or ["C2"]
dev = torch.device("cuda", 0)


def timed(fu, fi, users, csr):
    for _ in range(2):
        out = evaluate.topk_device(fu, fi, users, csr, 20, "bf16x3")
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = evaluate.topk_device(fu, fi, users, csr, 20, "bf16x3")
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return round(sorted(ts)[2], 3), out


tables = {}
for name in names:
    torch.manual_seed(0)
    sg = synth.make_graph(name)
    shp = synth.SHAPES[name]
    gr = graph.build_graph(sg.train_edges, sg.num_users, sg.num_items, sg.cred, shp["variant"], dev)
    d = shp["emb_dim"]
    eu = torch.nn.init.xavier_uniform_(torch.empty(sg.num_users, d)).to(dev)
    ei = torch.nn.init.xavier_uniform_(torch.empty(sg.num_items, d)).to(dev)
    fu, fi = model.propagate_forward(gr, eu, ei, shp["num_layers"], shp["order"])
    tables[name] = (fu, fi, torch.arange(sg.num_users, device=dev), (gr.samp_indptr, gr.samp_idx), sg.num_users,
                    sg.num_items, d)

if once:
    for fu, fi, users, csr, *_ in tables.values():
        for _ in range(2):
            evaluate.topk_device(fu, fi, users, csr, 20, "bf16x3")
    torch.cuda.synchronize()
elif variants:
    ref = {}
    for groups in (1, 2):
        _lib.set_option("EVAL_GROUPS", groups)
        for name, (fu, fi, users, csr, U, I, d) in tables.items():
            ms, (ids, sc) = timed(fu, fi, users, csr)
            line = {"workload": name, "groups": groups, "ms": ms}
            if name not in ref:
                ref[name] = (ids.clone(), sc.clone())
            else:
                line["ids_equal"] = bool(torch.equal(ids, ref[name][0]))
                line["score_bits_equal"] = bool(torch.equal(sc.view(torch.int32), ref[name][1].view(torch.int32)))
            print(json.dumps(line), flush=True)
    _lib.set_option("EVAL_GROUPS", 0)
else:
    out = {}
    for name, (fu, fi, users, csr, U, I, d) in tables.items():
        res = {}
        for opt in (0, 2, 6, 14, 0):
            _lib.set_option("EVAL_DEBUG", opt)
            res.setdefault(str(opt), []).append(timed(fu, fi, users, csr)[0])
        _lib.set_option("EVAL_DEBUG", 0)
        out[name] = {"users": U, "items": I, "d": d, "ms_by_option": res}
    print(json.dumps(out))
