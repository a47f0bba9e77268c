"""Recipe for `oracle/_ref/`: the UNMODIFIED reference scripts of the hot path, for the timed reference arm.

TEST / BENCH INFRASTRUCTURE ONLY.  `/root/reference` exists in the build container, not on the GPU box; the
reference is three stand-alone Python scripts (no package, no build), so "building" it means placing byte-identical
copies where the GPU box can import them:

    oracle/_ref/lightgcn_cu.py                   <- /root/reference/lightgcn_cu.py
    oracle/_ref/lighgcn_cu_pop.py                <- /root/reference/Version-2/lighgcn_cu_pop.py
    oracle/_ref/lightgcn_cu_pop_degree_aware.py  <- /root/reference/version_1/lightgcn_cu_pop_Degree-Aware Message.py
    oracle/_ref/MANIFEST.json                    md5 of every file + the versions it was staged with

`oracle/_ref/` is git-ignored (reference sources never enter this repository's history) but NOT gpurun-ignored, so it
travels with the snapshot.  `__graft_entry__.build()` runs this when /root/reference is present.  Only
`bench.py --impl reference` (through oracle/ref_runner.py) and tests import what it stages.
"""
from __future__ import annotations

import hashlib
import json
import pathlib
import shutil
import sys

HERE = pathlib.Path(__file__).resolve().parent
REF = pathlib.Path("/root/reference")
DEST = HERE / "_ref"
FILES = {
    "lightgcn_cu.py": REF / "lightgcn_cu.py",
    "lighgcn_cu_pop.py": REF / "Version-2" / "lighgcn_cu_pop.py",
    "lightgcn_cu_pop_degree_aware.py": REF / "version_1" / "lightgcn_cu_pop_Degree-Aware Message.py",
}


def md5(path: pathlib.Path) -> str:
    return hashlib.md5(path.read_bytes()).hexdigest()


def stage(verbose: bool = True) -> bool:
    """Copy the scripts (byte for byte) into oracle/_ref/.  Returns False when /root/reference is absent."""
    if not all(src.exists() for src in FILES.values()):
        if verbose:
            print(f"[make_ref] {REF} not present: nothing staged (the GPU box uses what travelled with the snapshot)")
        return False
    DEST.mkdir(exist_ok=True)
    manifest = {}
    for name, src in FILES.items():
        shutil.copyfile(src, DEST / name)
        assert md5(DEST / name) == md5(src)
        manifest[name] = {"source": str(src), "md5": md5(src), "bytes": src.stat().st_size}
    try:
        import numpy
        import torch
        manifest["_versions"] = {"torch": torch.__version__, "numpy": numpy.__version__,
                                 "python": sys.version.split()[0]}
    except Exception:       # noqa: BLE001
        pass
    (DEST / "MANIFEST.json").write_text(json.dumps(manifest, indent=1))
    if verbose:
        print(f"[make_ref] staged {len(FILES)} reference scripts in {DEST}")
    return True


if __name__ == "__main__":
    stage()
