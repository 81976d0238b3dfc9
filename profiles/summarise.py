#!/usr/bin/env python
"""Turns the gpurun_out/ ncu artefacts of one capture into the committed summaries under profiles/.

    python profiles/summarise.py <tag> <kernel regex name> "<title>"
reads gpurun_out/launches_<tag>.csv and gpurun_out/walk_<tag>.ncu-rep (via `ncu -i ... --page raw --csv`) and writes
profiles/launches_<tag>.md, profiles/kernel_<tag>.md and profiles/traffic.json (DRAM bytes per launch of the kernel).
"""
import collections
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = [
    'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sectors.sum', 'lts__t_sector_hit_rate.pct',
    'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
    'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__warps_eligible.avg.per_cycle_active',
    'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__inst_executed.sum', 'launch__registers_per_thread',
    'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic', 'sm__cycles_elapsed.max',
    'sm__cycles_active.avg', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
    'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
]


def launches(tag):
    path = os.path.join(ROOT, 'gpurun_out', f'launches_{tag}.csv')
    rows = [r for r in csv.reader(open(path)) if r and not r[0].startswith('==')]
    hdr, data = rows[0], rows[1:]
    ki, mi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    agg = collections.OrderedDict()
    for r in data:
        name = r[ki].split('(')[0]
        if 'cub' in name:
            name = 'cub::' + name.split('::')[-1].split('<')[0]
        if 'at::' in name:
            name = 'torch fill (L2 flush, outside the timed spans)'
        v = float(r[mi].replace(',', ''))
        u = r[ui]
        v *= {'ns': 1e-3, 'nsecond': 1e-3, 'us': 1, 'usecond': 1, 'ms': 1e3, 'msecond': 1e3, 's': 1e6, 'second': 1e6}[u]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    out = io.StringIO()
    out.write(f"# ncu launch list, capture {tag}\n\n")
    out.write("Command: `ncu --metrics gpu__time_duration.sum --clock-control none -c 260 --csv python bench.py --steps 2 "
              "--warmup 3 --cpu-reps 0` (1 x B200; cold-cache, serialised: compare SHARES, not absolutes). Microseconds.\n"
              "`fillFloat` is the map clear between steps and the torch fill is the L2 flush; both are outside the "
              "timed spans of bench.py.\n\n| kernel | launches | total us | mean us | share |\n|---|---:|---:|---:|---:|\n")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.write(f"| {k} | {n} | {t:.1f} | {t / n:.1f} | {100 * t / tot:.1f}% |\n")
    open(os.path.join(ROOT, 'profiles', f'launches_{tag}.md'), 'w').write(out.getvalue())
    return out.getvalue()


def kernel(tag, name, title, reading=""):
    rep = os.path.join(ROOT, 'gpurun_out', f'walk_{tag}.ncu-rep')
    raw = subprocess.check_output(['ncu', '-i', rep, '--page', 'raw', '--csv']).decode()
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, r = rows[0], rows[1], rows[2]
    lines = [f"# ncu --set full: {name}, capture {tag} ({title})\n\n",
             f"Command: `ncu --set full --clock-control none --import-source on -k regex:{name} -s 5 -c 1 python bench.py "
             f"--steps 2 --warmup 3 --cpu-reps 0` (1 x B200).\n\n| metric | value | unit |\n|---|---:|---|\n"]
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            lines.append(f"| {k} | {r[i]} | {units[i]} |\n")
    stalls = sorted(((float(r[hdr.index(h)]), h) for h in hdr
                     if 'average_warps_issue_stalled' in h and h.endswith('per_issue_active.ratio')), reverse=True)
    lines.append("\nTop stall reasons (warps per issue-active cycle):\n\n")
    for v, h in stalls[:6]:
        lines.append(f"* {h.split('issue_stalled_')[1].split('_per_issue')[0]}: {v:.2f}\n")
    if reading:
        lines.append("\n" + reading + "\n")
    open(os.path.join(ROOT, 'profiles', f'kernel_{tag}.md'), 'w').write(''.join(lines))
    # (profiles/traffic.json is written by tools/ncu_traffic.py, keyed by the hash of the CUDA sources)
    return ''.join(lines)


if __name__ == '__main__':
    tag, name, title = sys.argv[1], sys.argv[2], sys.argv[3]
    reading = sys.argv[4] if len(sys.argv) > 4 else ""
    print(launches(tag))
    print(kernel(tag, name, title, reading))
