"""Summarise ncu output into the text files kept under profiles/ (run in the build container).
    python profiles/summarize.py launches <launches.csv>
    python profiles/summarize.py raw <prof.ncu-rep>
    python profiles/summarize.py source <prof.ncu-rep>     (warp-specialised kernels: stall samples per warp role)
"""
import collections
import csv
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'launch__waves_per_multiprocessor', 'sm__maximum_warps_per_active_cycle_pct',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'smsp__cycles_active.avg',
        'sm__cycles_elapsed.max', 'smsp__inst_executed.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio']


def launches(path):
    lines = [l for l in open(path) if not l.startswith('==')]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        name = row['Kernel Name'].split('(')[0]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(row['Metric Value'])
    tot = sum(a[1] for a in agg.values())
    print(f"# {path}: {sum(a[0] for a in agg.values())} launches, {tot/1e3:.1f} us of kernel time (ncu: cold cache, serialised)")
    for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{t/1e3:10.1f} us  {c:5d} launches  {t/c/1e3:8.2f} us avg  {100*t/tot:5.1f} %  {k[:100]}")


def raw(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print('---- ' + r[idx['Kernel Name']][:110])
        for k in KEYS:
            if k in idx:
                print(f"  {k:84s} {r[idx[k]][:24]:>24s} {units[idx[k]]}")


def source(path):
    """Source page of a capture made with --import-source on: the SASS lines are cut into the code regions of a
    warp-specialised kernel at its role landmarks (first UTCHMMA = MMA issuer, the first 128-bit shared store after it =
    mask builders, first UTMALDG = producer, first LDTM = epilogue), and the warp-state
    samples of every region are summed by stall reason.  Inside the epilogue the samples are further split by how often
    a line executes per tile (scan = once per chunk, hit rounds / merge = data dependent)."""
    out = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    print('---- ' + rows[0][1][:110])
    hdr, data = rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    num = lambda r, k: int(r[ix[k]] or 0)
    src = [r[ix['Source']].strip() for r in data]
    stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    first = lambda tok: next(i for i, t in enumerate(src) if tok in t)
    back_to_branch = lambda i: max(k for k in range(i) if src[k].startswith('BRA ') or ' EXIT' in src[k]) + 1
    mma = first('UTCHMMA')
    mask = next(i for i in range(mma, len(src)) if src[i].startswith('STS.128'))     # the mask builders' store
    marks = sorted([(back_to_branch(mma), 'MMA issuer'), (back_to_branch(first('UTMALDG')), 'TMA producer'),
                    (back_to_branch(mask), 'mask builders'), (back_to_branch(first('LDTM')), 'epilogue (scan + lists)')])
    last_ldtm = max(i for i, t in enumerate(src) if 'LDTM' in t)
    end_epi = next(i for i in range(last_ldtm, len(src)) if 'BAR.SYNC' in src[i])
    regions = [(0, marks[0][0], 'prologue')]
    for (a, name), nxt in zip(marks, [m[0] for m in marks[1:]] + [end_epi]):
        regions.append((a, nxt, name))
    regions.append((end_epi, len(src), 'exit barrier + out-of-line spin loops'))
    total = sum(num(r, '# Samples') for r in data)
    print(f"  {total} warp-state samples, {sum(num(r, 'Instructions Executed') for r in data) / 1e6:.1f} M warp instructions")
    for a, b, name in regions:
        smp = sum(num(data[k], '# Samples') for k in range(a, b))
        agg = {h: sum(num(data[k], h) for k in range(a, b)) for h in stalls}
        top = ', '.join(f"{h[6:]} {100 * v / max(smp, 1):.0f} %" for h, v in sorted(agg.items(), key=lambda x: -x[1])[:5] if v)
        print(f"  {name:40s} lines {a:5d}-{b:5d}  {smp:6d} samples ({100 * smp / total:4.1f} %)  {top}")
    a, b = next((x, y) for x, y, n in regions if n.startswith('epilogue'))
    per_tile = num(data[first('LDTM')], 'Instructions Executed')
    buckets = collections.OrderedDict((k, [0, 0]) for k in ('once per tile (scan, waits)', 'data dependent (hit rounds, merge)'))
    for k in range(a, b):
        key = 'once per tile (scan, waits)' if num(data[k], 'Instructions Executed') == per_tile else \
              'data dependent (hit rounds, merge)'
        buckets[key][0] += num(data[k], '# Samples')
        buckets[key][1] += num(data[k], 'Instructions Executed')
    for key, (smp, ins) in buckets.items():
        print(f"    epilogue, {key:36s} {smp:6d} samples, {ins / max(per_tile, 1):7.1f} instructions per tile and warp")


if __name__ == '__main__':
    {'launches': launches, 'raw': raw, 'source': source}[sys.argv[1]](sys.argv[2])
