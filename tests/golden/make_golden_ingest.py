"""Golden fixture for the JSONL ingest: a tiny review file and what the reference writes for it.
    python tests/golden/make_golden_ingest.py     (build container only)"""
import importlib.util, io, contextlib, json, pathlib, pickle, tempfile
import numpy as np
ROOT = pathlib.Path(__file__).resolve().parents[2]
spec = importlib.util.spec_from_file_location("ref_cu", "/root/reference/lightgcn_cu.py")
ref = importlib.util.module_from_spec(spec); spec.loader.exec_module(ref)
rng = np.random.default_rng(3)
lines = []
for k in range(400):
    rec = {"user_id": f"U{rng.integers(0, 60):03d}", "parent_asin": f"B{rng.integers(0, 90):04d}",
           "rating": float(rng.integers(1, 6)), "text": "x" * int(rng.integers(0, 5))}
    if k % 37 == 0: rec.pop("user_id")
    if k % 41 == 0: rec["rating"] = "n/a"
    if k % 43 == 0: rec["rating"] = "5"
    lines.append(json.dumps(rec).encode())
lines[10] = b'{"broken json'
lines[20] = b""
lines[30] = json.dumps({"user_id": "U\xe9t\xe9", "parent_asin": "B0001", "rating": 5}).encode("latin-1")   # invalid utf-8
lines += lines[100:110]                                                                                       # duplicate pairs
jsonl = ROOT / "tests" / "golden" / "tiny_reviews.jsonl"
jsonl.write_bytes(b"\n".join(lines) + b"\n")
with tempfile.TemporaryDirectory() as td:
    ref.cfg.jsonl_path, ref.cfg.out_dir = str(jsonl), td
    with contextlib.redirect_stdout(io.StringIO()):
        ref.build_graph_from_jsonl()
    out = {k: np.load(pathlib.Path(td) / "npy" / f"{k}_edges.npy") for k in ("train", "val", "test")}
    u2 = pickle.load(open(pathlib.Path(td) / "model" / "user2idx.pkl", "rb"))
    i2 = pickle.load(open(pathlib.Path(td) / "model" / "item2idx.pkl", "rb"))
np.savez_compressed(ROOT / "tests" / "golden" / "tiny_ingest.npz", users=np.array(list(u2.keys())),
                    user_ids=np.array(list(u2.values())), items=np.array(list(i2.keys())),
                    item_ids=np.array(list(i2.values())), **out)
print({k: v.shape for k, v in out.items()}, len(u2), len(i2))
