"""Ingest: Amazon-review JSONL -> on-disk graph files, with the reference's formats and split.

    split_bucket            lightgcn_cu.py:107-116   md5(f"{uid}|{iid}") -> train / val / test
    iter_jsonl_records      lightgcn_cu.py:141-159   tolerant line reader
    build_graph_from_jsonl  lightgcn_cu.py:165-253   two streaming passes
Outputs (read back by train.train_lightgcn):
    <out_dir>/npy/{train,val,test}_edges.npy    int32 [2, E]
    <out_dir>/model/{user2idx,item2idx}.pkl     dict str -> int, ids in order of first appearance
Host-side I/O only: nothing here touches the GPU.
"""
from __future__ import annotations

import array
import hashlib
import json
import pickle
from pathlib import Path

import numpy as np

from . import config


def _as_float(x):
    try:
        return float(x)
    except Exception:
        return None


def is_positive_interaction(rec: dict, cfg=None) -> bool:
    cfg = cfg or config.cfg
    if rec.get(cfg.user_key) is None or rec.get(cfg.item_key) is None:
        return False
    rating = _as_float(rec.get(cfg.rating_key))
    return rating is not None and rating >= cfg.pos_rating_threshold


def split_bucket(uid: str, iid: str, cfg=None) -> str:
    cfg = cfg or config.cfg
    x = int(hashlib.md5(f"{uid}|{iid}".encode("utf-8")).hexdigest()[:8], 16) / 0xFFFFFFFF
    if x < cfg.train_p:
        return "train"
    return "val" if x < cfg.train_p + cfg.val_p else "test"


def ensure_paths(cfg=None):
    cfg = cfg or config.cfg
    p = Path(cfg.jsonl_path)
    if not p.exists():
        raise FileNotFoundError(f"JSONL not found:\n  {p}\n\nCurrent working dir: {Path.cwd()}\n"
                                "Set cfg.jsonl_path to your absolute JSONL path.\n")


def iter_jsonl_records(path: Path, cfg=None):
    cfg = cfg or config.cfg
    bad = 0
    with open(path, "rb") as f:
        for n, raw in enumerate(f, start=1):
            line = raw.decode("utf-8", errors=cfg.decode_errors).strip()
            if not line:
                continue
            try:
                yield n, json.loads(line)
            except json.JSONDecodeError:
                bad += 1
                if bad <= 5:
                    print(f"[WARN] Skipping invalid JSON at line {n}")
    if bad:
        print(f"[WARN] Total invalid JSON lines skipped: {bad:,}")


def build_graph_from_jsonl(cfg=None):
    cfg = cfg or config.cfg
    ensure_paths(cfg)
    out = Path(cfg.out_dir)
    (out / "model").mkdir(parents=True, exist_ok=True)
    (out / "npy").mkdir(parents=True, exist_ok=True)
    src = Path(cfg.jsonl_path)

    user2idx, item2idx = {}, {}
    # One pass suffices (ids and buckets are per record).  Edges are appended to flat int32 arrays, 8 bytes per edge
    # like the reference's preallocated [2, E] arrays (lightgcn_cu.py:211-237) -- a Python list of tuples would cost
    # ~100 bytes per edge, many GB on the tens of millions of positives of the real dataset.
    rows = {"train": array.array("i"), "val": array.array("i"), "test": array.array("i")}
    for n, rec in iter_jsonl_records(src, cfg):
        if not is_positive_interaction(rec, cfg):
            continue
        uid, iid = rec[cfg.user_key], rec[cfg.item_key]
        u = user2idx.setdefault(uid, len(user2idx))
        i = item2idx.setdefault(iid, len(item2idx))
        b = rows[split_bucket(uid, iid, cfg)]
        b.append(u)
        b.append(i)
        if n % cfg.print_every == 0:
            print(f"PASS {n:,} | users={len(user2idx):,} items={len(item2idx):,} "
                  + " ".join(f"{k}={len(v) // 2:,}" for k, v in rows.items()))
    print("Users:", len(user2idx), "Items:", len(item2idx), "Positive edges:", sum(len(v) // 2 for v in rows.values()))
    print("Split counts:", {k: len(v) // 2 for k, v in rows.items()})
    for name, obj in (("user2idx", user2idx), ("item2idx", item2idx)):
        with open(out / "model" / f"{name}.pkl", "wb") as f:
            pickle.dump(obj, f, protocol=pickle.HIGHEST_PROTOCOL)
    for k, v in rows.items():
        arr = np.frombuffer(v, dtype=np.int32).reshape(-1, 2).T if len(v) else np.empty((2, 0), dtype=np.int32)
        np.save(out / "npy" / f"{k}_edges.npy", np.ascontiguousarray(arr))
    print("\n✅ Saved graph files to:", out)
