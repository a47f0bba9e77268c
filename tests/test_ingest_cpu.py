"""Ingest (host I/O): same files as the reference writes for the same JSONL (fixture made by running
lightgcn_cu.py:165-253 on tests/golden/tiny_reviews.jsonl)."""
import pathlib
import pickle

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parents[1]


def test_build_graph_from_jsonl_matches_reference(tmp_path, capsys):
    from credgcn import config, ingest
    gold = np.load(ROOT / "tests" / "golden" / "tiny_ingest.npz")
    cfg = config.CFG()
    cfg.jsonl_path, cfg.out_dir = str(ROOT / "tests" / "golden" / "tiny_reviews.jsonl"), str(tmp_path)
    ingest.build_graph_from_jsonl(cfg)
    for k in ("train", "val", "test"):
        got = np.load(tmp_path / "npy" / f"{k}_edges.npy")
        assert got.dtype == np.int32 and got.shape[0] == 2
        np.testing.assert_array_equal(got, gold[k])
    u2 = pickle.load(open(tmp_path / "model" / "user2idx.pkl", "rb"))
    i2 = pickle.load(open(tmp_path / "model" / "item2idx.pkl", "rb"))
    assert list(u2.keys()) == gold["users"].tolist() and list(u2.values()) == gold["user_ids"].tolist()
    assert list(i2.keys()) == gold["items"].tolist() and list(i2.values()) == gold["item_ids"].tolist()
    assert "invalid JSON" in capsys.readouterr().out


def test_split_bucket_is_md5_of_the_pair():
    from credgcn import ingest
    seen = {ingest.split_bucket(f"u{k}", f"i{k % 7}") for k in range(300)}
    assert seen == {"train", "val", "test"}
    assert ingest.split_bucket("A", "B") == ingest.split_bucket("A", "B")


def test_credibility_csv_loader(tmp_path, capsys):
    from credgcn import train
    user2idx = {"a": 0, "b": 1, "c": 2, "d": 3}
    p = tmp_path / "cred.csv"
    p.write_text("user_id,user_idx,credibility\na,9,0.25\nzz,1,0.5\nc,2,1.7\nd,3,oops\n")
    cred = train.load_credibility_vector(str(p), user2idx)            # user_id wins over user_idx
    np.testing.assert_allclose(cred, [0.25, 1.0, 1.0, 1.0])
    p.write_text("user_idx,credibility\n1,0.125\n7,0.5\n-1,0.5\n")
    np.testing.assert_allclose(train.load_credibility_vector(str(p), user2idx), [1.0, 0.125, 1.0, 1.0])
    assert train.load_credibility_vector(str(tmp_path / "missing.csv"), user2idx).tolist() == [1.0] * 4
