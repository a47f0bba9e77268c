"""profiles/traffic.py <workload> <prof.ncu-rep> [n_layers]   (run in the build container)

DRAM bytes of the SpMM launches of ONE training step from an `ncu --set full` capture of bench.py's eager step
(`--profile-from-start off -k regex:k_spmm -c 20 ... --steps 1`): dram__bytes_read.sum + dram__bytes_write.sum of every
k_spmm / k_spmm_ring / k_spmm_finish launch.  Writes / updates profiles/r2_traffic.json, which bench.py reads for
roofline.traffic and roofline.frac (SURVEY 8d ii: for tables beyond L2 the DRAM bytes are the ones that bound the kernel).
A step has 4K product launches (K forward pairs + K adjoint pairs); when the capture holds fewer, the missing adjoint
launches are filled with the captured adjoint launch of the same kind and the entry says so.
"""
import csv
import io
import json
import pathlib
import subprocess
import sys


def launches(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    idx = {h: i for i, h in enumerate(rows[0])}
    res = []
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "")
        res.append(dict(kernel=name, ms=float(r[idx["gpu__time_duration.sum"]]),
                        gb=float(r[idx["dram__bytes_read.sum"]]) + float(r[idx["dram__bytes_write.sum"]]),
                        l2_hit=float(r[idx["lts__t_sector_hit_rate.pct"]])))
    return res


def main():
    name, rep = sys.argv[1], sys.argv[2]
    K = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    ls = launches(rep)
    prod = [l for l in ls if "finish" not in l["kernel"]]
    fin = [l for l in ls if "finish" in l["kernel"]]
    note = f"all {4 * K} product launches of the step captured"
    if len(prod) < 4 * K:
        # forward = first 2K (item, user pairs); adjoint = NZ item launch, then user / item alternating
        adj = prod[2 * K:]
        users = [l for i, l in enumerate(adj) if i % 2 == 1]
        items = [l for i, l in enumerate(adj) if i % 2 == 0 and i > 0]
        filled = list(prod)
        while len(filled) < 4 * K:
            i = len(filled) - 2 * K
            filled.append(dict((users if i % 2 == 1 else items)[-1], filled=True))
        note = (f"{len(prod)} of {4 * K} product launches captured; the missing adjoint launches are filled with the "
                "captured adjoint launch of the same kind")
        prod = filled
    n_fin = len(fin) if len(fin) >= 2 * K else 2 * K * (1 if fin else 0)
    total = sum(l["gb"] for l in prod) + (sum(l["gb"] for l in fin) / max(len(fin), 1)) * n_fin
    path = pathlib.Path(__file__).resolve().parent / "r2_traffic.json"
    data = json.loads(path.read_text()) if path.exists() else {
        "_comment": "DRAM traffic of the SpMM launches of one training step (dram__bytes_read.sum + "
                    "dram__bytes_write.sum, ncu --set full --clock-control none); written by profiles/traffic.py, read "
                    "by bench.py for roofline.traffic / roofline.frac"}
    data[name] = {"bytes_per_step": total * 1e9, "launches_per_step": 4 * K, "source": f"profiles/{pathlib.Path(rep).stem}"
                  ".txt (ncu --set full of bench.py's eager step)", "note": note,
                  "launches": [dict(kernel=l["kernel"], ms=round(l["ms"], 3), gb=round(l["gb"], 2),
                                    l2_hit_pct=round(l["l2_hit"], 1), **({"filled": True} if l.get("filled") else {}))
                               for l in prod]}
    path.write_text(json.dumps(data, indent=1))
    print(f"{name}: {total:.1f} GB per step over {4 * K} launches ({note})")


if __name__ == "__main__":
    main()
