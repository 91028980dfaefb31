#!/usr/bin/env python
"""Where is the GPU idle inside a converged benchmark sweep?  Kernel timeline (torch.profiler / CUPTI activity records: all
kernels of the process, also those of libb200npc.so) of one sweep after warm-up; gaps between consecutive device
activities, aggregated by (activity before the gap -> activity after it).

    python profiles/timeline_gaps.py [L=40] [chi=1024] [warm=4]
"""
import json
import os
import re
import sys
import tempfile

import torch
from torch.profiler import profile, ProfilerActivity

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from tenpy_b200 import backend  # noqa: E402
from tenpy_b200.algorithms import dmrg  # noqa: E402
from tenpy_b200.models import TFIChain  # noqa: E402


def short(name):
    name = name.replace('(anonymous namespace)::', '').replace('<unnamed>::', '')
    name = re.sub(r'\(.*', '', name)
    name = re.sub(r'<.*', '', name)
    return name.split('::')[-1].strip()[:40] or 'anonymous'


def main():
    L = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    chi = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    warm = int(sys.argv[3]) if len(sys.argv) > 3 else 4
    backend.get_lib()
    model = TFIChain({'L': L, 'J': 1., 'g': 1., 'conserve': None})
    psi = bench.synthetic_mps(model, L, chi, 2, seed=0)
    eng = dmrg.TwoSiteDMRGEngine(psi, model, {
        'mixer': None, 'combine': True, 'diag_method': 'lanczos', 'svd_warm_start': False,
        'trunc_params': {'chi_max': chi, 'svd_min': 1e-45, 'trunc_cut': None, 'svd_deflation_tol': 1e-10},
        'lanczos_params': {'N_min': 10, 'N_max': 10}})
    for _ in range(warm):
        eng.sweep()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    eng.sweep()
    ev1.record()
    torch.cuda.synchronize()
    plain = ev0.elapsed_time(ev1) / 1e3
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        eng.sweep()
        torch.cuda.synchronize()
    path = os.path.join(tempfile.mkdtemp(), 'trace.json')
    prof.export_chrome_trace(path)
    with open(path) as f:
        tr = json.load(f)
    evs = [e for e in tr['traceEvents'] if e.get('ph') == 'X' and e.get('cat') in ('kernel', 'gpu_memcpy', 'gpu_memset')]
    evs.sort(key=lambda e: e['ts'])
    busy = sum(e['dur'] for e in evs)
    span = evs[-1]['ts'] + evs[-1]['dur'] - evs[0]['ts']
    gaps = {}
    hist = {'<5us': [0, 0.], '5-20us': [0, 0.], '20-100us': [0, 0.], '100us-1ms': [0, 0.], '>1ms': [0, 0.]}
    end = evs[0]['ts'] + evs[0]['dur']
    prev = evs[0]
    for e in evs[1:]:
        g = e['ts'] - end
        if g > 0:
            key = short(prev['name']) + ' -> ' + short(e['name'])
            c = gaps.setdefault(key, [0, 0.])
            c[0] += 1
            c[1] += g
            b = '<5us' if g < 5 else '5-20us' if g < 20 else '20-100us' if g < 100 else '100us-1ms' if g < 1000 else '>1ms'
            hist[b][0] += 1
            hist[b][1] += g
        if e['ts'] + e['dur'] > end:
            end = e['ts'] + e['dur']
            prev = e
    by_kernel = {}
    for e in evs:
        c = by_kernel.setdefault(short(e['name']), [0, 0.])
        c[0] += 1
        c[1] += e['dur']
    top = sorted(gaps.items(), key=lambda kv: -kv[1][1])[:40]
    out = {'L': L, 'chi': chi, 'sweep_s_plain': plain, 'profiled_span_s': span / 1e6, 'device_busy_s': busy / 1e6,
           'idle_s': (span - busy) / 1e6, 'activities': len(evs),
           'gap_histogram': {k: {'n': v[0], 'ms': round(v[1] / 1e3, 2)} for k, v in hist.items()},
           'top_gaps': [{'between': k, 'n': v[0], 'ms': round(v[1] / 1e3, 2), 'avg_us': round(v[1] / v[0], 1)} for k, v in top],
           'busy_by_kernel_ms': {k: [v[0], round(v[1] / 1e3, 2)] for k, v in sorted(by_kernel.items(), key=lambda kv: -kv[1][1])[:30]}}
    print(json.dumps(out))


if __name__ == '__main__':
    main()
