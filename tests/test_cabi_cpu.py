"""CPU-side checks of the drop-in boundary: the shared library loads, exports every symbol that
include/credgcn.h declares, and the product path refuses to run without CUDA (no CPU fallback).
No compute call is made here."""
import ctypes
import pathlib
import re

import numpy as np
import pytest
import torch

ROOT = pathlib.Path(__file__).resolve().parents[1]


def declared_symbols():
    text = (ROOT / "include" / "credgcn.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cgx_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from credgcn import _lib
    assert _lib.LIB_PATH.exists(), "build the extension first: python -c 'import __graft_entry__ as g; g.build()'"
    handle = ctypes.CDLL(str(_lib.LIB_PATH))
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(handle, n), f"{n} declared in include/credgcn.h but not exported"
    assert set(names) == set(_lib.EXPORTED), "ctypes signature table and header disagree"


def test_version_and_dim_table():
    from credgcn import _lib
    lib = _lib.lib()
    assert lib.cgx_version() >= 100
    assert [d for d in (8, 16, 32, 48, 64, 128, 256, 512) if lib.cgx_emb_dim_supported(d)] == [16, 32, 64, 128, 256]
    assert lib.cgx_graph_build_workspace_bytes(1000, 10, 10) > 3 * 8 * 1000


def test_options_table_and_eval_kernel_selection():
    """Every option of the header's enum is reachable by name from the binding, and the shape -> kernel decision of
    the evaluation (a host-side function, no launch) holds under both scanning-group settings: d <= 128 runs on the
    tensor cores with BF16X3, d = 256 only with the single-pass BF16 operand (shared-memory fit)."""
    from credgcn import _lib
    text = (ROOT / "include" / "credgcn.h").read_text()
    enum = dict(re.findall(r"^\s+CGX_OPT_([A-Z0-9_]+?)_? = (\d+)", text, flags=re.M))      # the enum's own lines
    count = int(enum.pop("COUNT"))
    assert {k: int(v) for k, v in enum.items()} == _lib.OPTIONS and count == len(_lib.OPTIONS)
    lib = _lib.lib()
    assert lib.cgx_get_option(count) == -1
    keep = _lib.get_option("EVAL_GROUPS")
    try:
        for groups in (0, 1, 2):
            _lib.set_option("EVAL_GROUPS", groups)
            x3, b = _lib.PRECISIONS["bf16x3"], _lib.PRECISIONS["bf16"]
            assert [lib.cgx_eval_topk_uses_tensor_cores(d, 20, x3) for d in (16, 64, 128, 256)] == [1, 1, 1, 0]
            assert lib.cgx_eval_topk_uses_tensor_cores(256, 20, b) == 1
            assert lib.cgx_eval_topk_uses_tensor_cores(64, 40, x3) == 1 and lib.cgx_eval_topk_uses_tensor_cores(64, 60, x3) == 0
            assert lib.cgx_eval_topk_uses_tensor_cores(64, 20, _lib.PRECISIONS["fp32"]) == 0
    finally:
        _lib.set_option("EVAL_GROUPS", keep)


def test_sass_carries_the_blackwell_instructions():
    """The shipped library is sm_100a code that uses what DESIGN.md says it uses: tcgen05 MMAs with TMEM loads and
    commits, TMA tile loads, 256-bit gathers with L2 eviction priorities, the in-switch reduction of the NVLS exchange,
    programmatic dependent launch (profiles/sass_evidence.py; profiles/r2_sass_evidence.txt holds the counts)."""
    import shutil
    import sys
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    sys.path.insert(0, str(ROOT / "profiles"))
    try:
        import sass_evidence
    finally:
        sys.path.pop(0)
    from credgcn import _lib
    got = sass_evidence.counts(_lib.LIB_PATH)
    assert all(n > 0 for n in got.values()), got


def test_struct_layout_matches_header():
    from credgcn._lib import CsrStruct
    # int32,int32,int64, 5 pointers (indptr, idx, val_fwd, val_bwd, perm), int32,int32, 2 pointers, int32,int32,
    # 3 pointers (arrive, work, idx_hint)
    assert ctypes.sizeof(CsrStruct) == 4 + 4 + 8 + 5 * 8 + 4 + 4 + 2 * 8 + 4 + 4 + 8 + 8 + 8
    assert CsrStruct.work.offset == 96 and CsrStruct.idx_hint.offset == 104
    assert CsrStruct.n_huge.offset == 80 and CsrStruct.arrive.offset == 88
    assert CsrStruct.perm.offset == 48 and CsrStruct.n_long.offset == 56 and CsrStruct.chunk_ptr.offset == 64


def test_no_cpu_fallback():
    from credgcn import _lib, graph
    edges = np.array([[0, 1], [1, 0]], dtype=np.int32)
    with pytest.raises(_lib.CgxError):
        graph.build_graph(edges, 2, 2, np.ones(2, np.float32), "v2", device="cpu")
    with pytest.raises(_lib.CgxError):
        _lib.ptr(torch.zeros(4))


def test_product_does_not_import_oracle():
    pkg = ROOT / "beyond-binary-fake-user-detection-a-credibility-aware-graph-based-recommender-system_b200"
    for f in pkg.glob("*.py"):
        src = f.read_text()
        assert "credgcn_oracle" not in src and "import oracle" not in src, f"{f.name} touches the oracle"


def test_config_mirrors_reference_fields():
    from credgcn.config import CFG
    c = CFG()
    assert (c.emb_dim, c.num_layers, c.lr, c.batch_size, c.epochs, c.Ks) == (64, 3, 1e-3, 4096, 400, (10, 20))
    assert (c.lambda_reg, c.reg, c.lambda_fair) == (1e-4, 1e-4, 0.0)
    assert (c.neg_mix_pop, c.neg_pop_gamma, c.neg_max_tries, c.cred_group_pct) == (0.7, 0.75, 50, 0.20)
    assert (c.sampled_negatives, c.seed, c.eval_mode) == (99, 42, "sampled")


def test_host_metrics_match_oracle(golden):
    """The vectorised NumPy metrics (host logic) against the oracle's per-user loop."""
    import credgcn_oracle as orc
    from credgcn import evaluate
    g = golden
    if "full_ranked" not in g:
        pytest.skip("no full-rank fixture for lightgcn_cu.py")
    te = (g["test_indptr"], g["test_indices"])
    users = np.flatnonzero(np.diff(te[0]) > 0)
    extra = g["tag"] == "v2"
    got = evaluate.metrics_from_ranked(g["full_ranked"], users, te, int(g["num_items"]), (10, 20), "full",
                                       g["item_pop"] if extra else None, int(g["total_train"]) if extra else 0,
                                       g["cred"] if extra else None)
    for K in (10, 20):
        want = g[f"full_{K}"]
        np.testing.assert_allclose([got[K]["precision"], got[K]["recall"], got[K]["ndcg"]], want[:3], rtol=1e-6)
        if extra:
            np.testing.assert_allclose(
                [got[K][k] for k in ("item_coverage", "avg_log_popularity", "avg_self_information", "cred_utility",
                                     "high_cred_recall", "low_cred_recall")], want[3:], rtol=1e-6)


def test_sampled_candidates_match_reference_stream(golden):
    from credgcn import evaluate
    g = golden
    tr = (g["csr_indptr"], g["csr_indices"])
    te = (g["test_indptr"], g["test_indices"])
    users, cands = evaluate.sampled_candidates(tr, te, int(g["num_items"]), 99, 42)
    # the reference ranked exactly these candidates: same set per user
    np.testing.assert_array_equal(np.sort(cands, 1)[:, :0].shape[0], g["sampled_ranked"].shape[0])
    for r in range(len(users)):
        assert set(g["sampled_ranked"][r].tolist()) <= set(cands[r].tolist())


def test_plain_c_caller_links_and_runs(tmp_path):
    """tests/c/abi_smoke.c is compiled with gcc against include/credgcn.h and linked to libcredgcn.so: the boundary
    is a C ABI (no C++ / torch types), usable from any host language's FFI.  Host-side entry points only."""
    import shutil
    import subprocess
    from credgcn import _lib
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    exe = tmp_path / "abi_smoke"
    lib_dir = _lib.LIB_PATH.parent
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", f"-I{ROOT / 'include'}", str(ROOT / "tests" / "c" / "abi_smoke.c"),
                    "-o", str(exe), f"-L{lib_dir}", "-lcredgcn", f"-Wl,-rpath,{lib_dir}"], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0 and "0 failure(s)" in out.stdout, out.stdout + out.stderr
