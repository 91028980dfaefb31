#!/usr/bin/env python
"""Summarise the ncu outputs of profiles/capture.sh into small committed files under profiles/."""
import csv
import os
import subprocess
import sys
from collections import defaultdict

OUT = os.path.dirname(os.path.abspath(__file__))
GO = os.path.join(os.path.dirname(OUT), 'gpurun_out')


def launches(tag='r01', what=None, command=None):
    path = os.path.join(GO, tag + '_launches.csv')
    rows = []
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    rd = csv.reader(lines)
    hdr = next(rd)
    ik, iv = hdr.index('Kernel Name'), hdr.index('Metric Value')
    for r in rd:
        if len(r) > iv:
            try:
                rows.append((r[ik], float(r[iv].replace(',', ''))))
            except ValueError:
                pass
    tot = defaultdict(lambda: [0, 0.])
    for name, ns in rows:
        key = name.split('(')[0]
        tot[key][0] += 1
        tot[key][1] += ns
    total = sum(v[1] for v in tot.values())
    with open(os.path.join(OUT, tag + '_launch_shares.md'), 'w') as f:
        f.write('# %s: per-kernel device time of %s (ncu gpu__time_duration.sum, serialised, '
                'cold cache: compare SHARES)\n\n' % (tag, what or 'the timed bench step'))
        f.write('command: `%s`\n\n' % (command or 'ncu --metrics gpu__time_duration.sum --clock-control none '
                                        '--profile-from-start off python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu'))
        f.write('| kernel | launches | total ms | share | avg us |\n|---|---:|---:|---:|---:|\n')
        for k, (n, ns) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            f.write('| `%s` | %d | %.1f | %.3f | %.1f |\n' % (k, n, ns / 1e6, ns / total, ns / n / 1e3))
        f.write('\n%d launches, %.1f ms total\n' % (len(rows), total / 1e6))
    print('wrote launch shares:', len(rows), 'launches')


KEYS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.sum', 'smsp__inst_executed.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__warp_issue_stalled_barrier_per_warp_active.pct', 'smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct',
        'smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct', 'smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct',
        'smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct', 'smsp__warp_issue_stalled_wait_per_warp_active.pct',
        'sm__cycles_elapsed.max', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']


def full(tag, name):
    rep = os.path.join(GO, '%s_%s.ncu-rep' % (tag, name))
    if not os.path.exists(rep):
        print('missing', rep)
        return
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[0]
    with open(os.path.join(OUT, '%s_%s_ncu.md' % (tag, name)), 'w') as f:
        f.write('# %s: `ncu --set full --clock-control none` of %s (see profiles/capture.sh)\n\n' % (tag, name))
        for r in rows[2:]:
            f.write('## %s\n\n| metric | value |\n|---|---|\n' % r[hdr.index('Kernel Name')][:120])
            for k in KEYS:
                if k in hdr:
                    f.write('| %s | %s |\n' % (k, r[hdr.index(k)]))
            f.write('\n')
    print('wrote', name)


if __name__ == '__main__':
    tag = sys.argv[1] if len(sys.argv) > 1 else 'r01'
    if tag == 'r01d':
        launches(tag, 'ONE centre-bond update at chi=1024 (10 Lanczos matvecs in the split order, block SVD, '
                      'environment update) after 3 warm-up sweeps',
                 'ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off '
                 'python profiles/bond_probe.py --bonds 1  (profiles/capture_r01d.sh)')
        full(tag, 'bond')
    else:
        launches(tag)
        full(tag, 'gemm')
        full(tag, 'jacobi')
