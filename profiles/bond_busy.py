#!/usr/bin/env python
"""Per phase of a bond update of the benchmark sweep: wall clock (device synchronised at the phase boundaries), host time
until the phase's last launch was issued, and the CUDA-event time of the library calls inside it.  wall - events = the
part of the phase in which the GPU waits for the host.  Bonds at full chi and the (launch-bound) edge bonds separately.

    python profiles/bond_busy.py [L=100] [chi=1024]
"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from tenpy_b200 import backend  # noqa: E402
from tenpy_b200.algorithms import dmrg  # noqa: E402
from tenpy_b200.models import TFIChain  # noqa: E402


def main():
    L = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    chi = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    lib = backend.get_lib()
    model = TFIChain({'L': L, 'J': 1., 'g': 1., 'conserve': None})
    psi = bench.synthetic_mps(model, L, chi, 2, seed=0)
    eng = dmrg.TwoSiteDMRGEngine(psi, model, {
        'mixer': None, 'combine': True, 'diag_method': 'lanczos', 'svd_warm_start': False,
        'trunc_params': {'chi_max': chi, 'svd_min': 1e-45, 'trunc_cut': None, 'svd_deflation_tol': 1e-10},
        'lanczos_params': {'N_min': 10, 'N_max': 10}})
    for _ in range(2):
        eng.sweep()
    acc = {}
    state = {'full': False}

    def ev_ms():
        return {k: v[1] for k, v in lib.profile_summary().items()}

    def timed(name, fn):
        def wrapper(*a, **k):
            torch.cuda.synchronize()
            lib.profile = {}
            t0 = time.perf_counter()
            r = fn(*a, **k)
            t1 = time.perf_counter()
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            fam = ev_ms()
            lib.profile = None
            key = ('full ' if state['full'] else 'edge ') + name
            d = acc.setdefault(key, {'n': 0, 'wall': 0., 'host_issue': 0., 'events': 0., 'fam': {}})
            d['n'] += 1
            d['wall'] += t2 - t0
            d['host_issue'] += t1 - t0
            d['events'] += sum(fam.values()) * 1e-3
            for f, ms in fam.items():
                d['fam'][f] = d['fam'].get(f, 0.) + ms
            return r
        return wrapper

    orig_prepare = eng.prepare_update_local

    def prepare():
        r = orig_prepare()
        i0 = eng.i0
        state['full'] = min(psi.get_B(i0, form=None).shape[0], psi.get_B(i0 + 1, form=None).shape[2]) >= chi
        return r
    eng.prepare_update_local = timed('prepare_update_local', prepare)
    eng.diag = timed('diag', eng.diag)
    eng.mixed_svd = timed('mixed_svd', eng.mixed_svd)
    eng.update_env = timed('update_env', eng.update_env)
    eng.set_B = timed('set_B', eng.set_B)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    eng.sweep()
    torch.cuda.synchronize()
    total = time.perf_counter() - t0
    out = {'L': L, 'chi': chi, 'sweep_s_with_phase_syncs': total, 'phases': {}}
    for k, d in sorted(acc.items()):
        out['phases'][k] = {'calls': d['n'], 'wall_ms_per_call': round(d['wall'] / d['n'] * 1e3, 3),
                            'host_issue_ms_per_call': round(d['host_issue'] / d['n'] * 1e3, 3),
                            'event_ms_per_call': round(d['events'] / d['n'] * 1e3, 3),
                            'wall_s_total': round(d['wall'], 3), 'gpu_waits_s_total': round(d['wall'] - d['events'], 3),
                            'families_ms_per_call': {f: round(ms / d['n'], 3) for f, ms in d['fam'].items()}}
    print(json.dumps(out))


if __name__ == '__main__':
    main()
