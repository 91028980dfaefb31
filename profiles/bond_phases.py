#!/usr/bin/env python
"""Where does a bond update of the benchmark spend its time?  Wall clock per phase of `TwoSiteDMRGEngine.update_local`
with a device synchronisation at every phase boundary (so host and device time of a phase add up; the sum is an upper
bound of the asynchronous sweep), for the identity-environment shortcut on and off.

    python profiles/bond_phases.py [L=30] [chi=1024]
"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from tenpy_b200 import backend  # noqa: E402
from tenpy_b200.algorithms import dmrg, mps_common  # noqa: E402
from tenpy_b200.linalg import krylov_based  # noqa: E402
from tenpy_b200.models import TFIChain  # noqa: E402


def main():
    L = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    chi = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    lib = backend.get_lib()
    model = TFIChain({'L': L, 'J': 1., 'g': 1., 'conserve': None})
    for identity in (True, False):
        psi = bench.synthetic_mps(model, L, chi, 2, seed=0)
        eng = dmrg.TwoSiteDMRGEngine(psi, model, {
            'mixer': None, 'combine': True, 'diag_method': 'lanczos', 'svd_warm_start': False, 'identity_env': identity,
            'trunc_params': {'chi_max': chi, 'svd_min': 1e-45, 'trunc_cut': None, 'svd_deflation_tol': 1e-10},
            'lanczos_params': {'N_min': 10, 'N_max': 10}})
        for _ in range(2):
            eng.sweep()
        acc = {}

        def timed(name, fn):
            def wrapper(*a, **k):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                r = fn(*a, **k)
                torch.cuda.synchronize()
                acc[name] = acc.get(name, 0.) + time.perf_counter() - t0
                return r
            return wrapper
        eng.prepare_update_local = timed('prepare_update_local', eng.prepare_update_local)
        eng.diag = timed('diag (Lanczos incl. matvecs)', eng.diag)
        eng.mixed_svd = timed('mixed_svd (svd_theta)', eng.mixed_svd)
        eng.update_env = timed('update_env', eng.update_env)
        eng.set_B = timed('set_B', eng.set_B)
        orig_setup = mps_common.TwoSiteH._identity_env_setup
        mps_common.TwoSiteH._identity_env_setup = timed('  of diag: identity_env_setup', orig_setup)
        orig_mv = mps_common.TwoSiteH.matvec
        mps_common.TwoSiteH.matvec = timed('  of diag: matvec', orig_mv)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        eng.sweep()
        torch.cuda.synchronize()
        total_sync = time.perf_counter() - t0
        mps_common.TwoSiteH._identity_env_setup = orig_setup
        mps_common.TwoSiteH.matvec = orig_mv
        for k in ('prepare_update_local', 'diag', 'mixed_svd', 'update_env', 'set_B'):
            eng.__dict__.pop(k, None)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        eng.sweep()
        torch.cuda.synchronize()
        total_async = time.perf_counter() - t0
        nb = 2 * (L - 2)
        print(json.dumps({'identity_env': identity, 'L': L, 'chi': chi, 'bonds': nb, 'sweep_s_with_phase_syncs': total_sync,
                          'sweep_s_async': total_async, 'ms_per_bond_by_phase': {k: round(v / nb * 1e3, 3) for k, v in acc.items()}}))


if __name__ == '__main__':
    main()
