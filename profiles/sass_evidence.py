"""Counts of the sm_100a instructions the design relies on, in the SASS of the shipped libcredgcn.so:
    python profiles/sass_evidence.py [path/to/libcredgcn.so]        (needs cuobjdump; no GPU)
tests/test_cabi_cpu.py::test_sass_carries_the_blackwell_instructions asserts the same list."""
import collections
import pathlib
import re
import subprocess
import sys

ROOT = pathlib.Path(__file__).resolve().parents[1]
# mnemonic prefix -> what it is evidence of
EVIDENCE = collections.OrderedDict([
    ("UTCHMMA", "tcgen05.mma (evaluation score tiles, accumulators in TMEM)"),
    ("LDTM", "tcgen05.ld (TMEM -> registers in the evaluation epilogue)"),
    ("UTCBAR", "tcgen05.commit onto mbarriers"),
    ("UTMALDG", "TMA tile loads (cp.async.bulk.tensor) of the item operand"),
    ("ELECT", "elect.sync (leader of the MMA issue) and ptxas' own single-lane selections"),
    ("SYNCS.PHASECHK", "mbarrier try_wait"),
    ("LDG.E.NA.ELL2.256", "256-bit gathers of HOT rows, L2 evict_last (SpMM, tables beyond L2)"),
    ("LDG.E.NA.EFL2.256", "256-bit gathers of cold rows / running sums, L2 evict_first"),
    ("STG.E.EFL2.256", "256-bit streaming stores of both SpMM outputs, L2 evict_first"),
    ("LDGMC.E.ADD.F32", "multimem.ld_reduce: the item-table sum inside the NVSwitch (NVLS exchange)"),
    ("ACQBULK", "griddepcontrol.wait (programmatic dependent launch between consecutive SpMMs)"),
    ("PREEXIT", "griddepcontrol.launch_dependents"),
])


def counts(lib):
    sass = subprocess.run(["cuobjdump", "-sass", str(lib)], capture_output=True, text=True, check=True).stdout
    ops = collections.Counter(m.group(1) for m in re.finditer(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d\s+)?([A-Z][A-Z0-9_.]*)", sass, re.M))
    return collections.OrderedDict((k, sum(v for op, v in ops.items() if op.startswith(k))) for k in EVIDENCE)


if __name__ == "__main__":
    lib = pathlib.Path(sys.argv[1]) if len(sys.argv) > 1 else next(ROOT.glob("beyond-binary-*_b200/libcredgcn.so"))
    for k, n in counts(lib).items():
        print(f"{n:6d}  {k:20s} {EVIDENCE[k]}")
