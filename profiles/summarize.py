"""Summarise ncu output into the text files kept under profiles/ (run in the build container).
    python profiles/summarize.py launches <launches.csv>
    python profiles/summarize.py raw <prof.ncu-rep>
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


if __name__ == '__main__':
    {'launches': launches, 'raw': raw}[sys.argv[1]](sys.argv[2])
