#!/usr/bin/env python
"""Summarise an `ncu --set full` report into a small committed markdown table (one row per captured kernel launch class):
duration, DRAM bytes / throughput, tensor-pipe activity, L2 throughput, occupancy, registers.

    python profiles/summarize_ncu.py gpurun_out/r02j_bond.ncu-rep profiles/r02j_bond_ncu.md "title"
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict

METRICS = OrderedDict([
    ('gpu__time_duration.sum', ('us', 1e-3)),
    ('dram__bytes_read.sum', ('MB read', 1.)),
    ('dram__bytes_write.sum', ('MB written', 1.)),
    ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', ('DRAM %', 1.)),
    ('lts__throughput.avg.pct_of_peak_sustained_elapsed', ('L2 %', 1.)),
    ('sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed', ('tensor pipe % (all)', 1.)),
    ('sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', ('TMEM busy %', 1.)),
    ('smsp__pipe_tensor_subpipe_dmma_cycles_active.avg', ('DMMA cycles', 1.)),
    ('sm__throughput.avg.pct_of_peak_sustained_elapsed', ('SM %', 1.)),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', ('occupancy %', 1.)),
    ('launch__registers_per_thread', ('regs', 1.)),
    ('launch__grid_size', ('grid', 1.)),
])


def main():
    rep, out, title = sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else ''
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ik = hdr.index('Kernel Name')
    cols = []
    for m in METRICS:
        hit = [i for i, h in enumerate(hdr) if h == m or h.endswith('.' + m)]
        if hit:
            cols.append((m, hit[0]))
    agg = OrderedDict()
    for r in rows[2:]:
        if len(r) <= ik:
            continue
        name = r[ik].split('(')[0].replace('b200::', '').replace('<unnamed>::', '')
        vals = []
        for m, i in cols:
            try:
                v = float(r[i].replace(',', ''))
            except ValueError:
                v = float('nan')
            u = units[i]
            if m.startswith('dram__bytes'):
                v = v * {'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1., 'Gbyte': 1e3}.get(u, 1.)
            if m == 'gpu__time_duration.sum':
                v = v * {'ns': 1e-3, 'us': 1., 'ms': 1e3, 'nsecond': 1e-3, 'usecond': 1., 'msecond': 1e3}.get(u, 1e-3)
            vals.append(v)
        agg.setdefault(name, []).append(vals)
    with open(out, 'w') as f:
        f.write('# %s\n\nsource: `%s` (ncu --set full --clock-control none; mean over the captured launches of each kernel; '
                'durations under the profiler are cold-cache and serialised)\n\n' % (title, rep))
        f.write('| kernel | n | ' + ' | '.join(METRICS[m][0] for m, _ in cols) + ' |\n')
        f.write('|---|---:|' + '---:|' * len(cols) + '\n')
        for name, lst in agg.items():
            n = len(lst)
            means = [sum(v[j] for v in lst) / n for j in range(len(cols))]
            f.write('| `%s` | %d | ' % (name, n) + ' | '.join('%.4g' % x for x in means) + ' |\n')
    print(open(out).read())


if __name__ == '__main__':
    main()
