#!/usr/bin/env python
"""The unit of the benchmark step under the profiler: `--bonds` two-site updates at the centre of a chi=1024 chain.

A full L=100 sweep is 196 such updates (~76 000 launches), far too many to replay under ncu; one centre bond is
the dominant unit (158 of the 196 bonds run at the full bond dimension).  A short chain (L=22 is the shortest with
chi=1024 at its centre) is swept once as warm-up, then `cudaProfilerStart/Stop` brackets the centre updates of the
next right-moving half sweep, so `ncu --profile-from-start off` sees exactly those launches.
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from tenpy_b200 import backend  # noqa: E402
from tenpy_b200.algorithms import dmrg  # noqa: E402
from tenpy_b200.models import TFIChain  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--L', type=int, default=22)
ap.add_argument('--chi', type=int, default=1024)
ap.add_argument('--bonds', type=int, default=1)
ap.add_argument('--lanczos-N', type=int, default=10)
ap.add_argument('--warm-sweeps', type=int, default=3, help='as bench.py: 3 untimed warm-up sweeps')
args = ap.parse_args()
torch.cuda.set_device(0)
lib = backend.get_lib()
L, chi = args.L, args.chi
model = TFIChain({'L': L, 'J': 1., 'g': 1., 'conserve': None})
psi = bench.synthetic_mps(model, L, chi, 2, seed=0)
eng = dmrg.TwoSiteDMRGEngine(psi, model, {
    'mixer': None, 'combine': True, 'diag_method': 'lanczos', 'svd_warm_start': False,
    'trunc_params': {'chi_max': chi, 'svd_min': 1e-45, 'trunc_cut': None, 'svd_deflation_tol': 1e-10},   # as bench.py
    'lanczos_params': {'N_min': args.lanczos_N, 'N_max': args.lanczos_N}})
for _ in range(args.warm_sweeps):
    eng.sweep()
torch.cuda.synchronize()
first = L // 2 - 1 - (args.bonds - 1) // 2
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
eng.E_trunc_list, eng.trunc_err_list = [], []
n0 = 0
for i0, move_right, upd in eng.get_sweep_schedule():
    if not move_right:
        break
    if i0 == first:
        torch.cuda.synchronize()
        lib.kernel_launch_count(reset=True)
        torch.cuda.profiler.start()
        ev0.record()
    eng.i0, eng.move_right, eng.update_LP_RP = i0, move_right, upd
    theta = eng.prepare_update_local()
    data = eng.update_local(theta)
    eng.update_env(**data)
    eng.post_update_local(**data)
    eng.free_no_longer_needed_envs()
    if i0 == first + args.bonds - 1:
        ev1.record()
        torch.cuda.profiler.stop()
        torch.cuda.synchronize()
        n0 = lib.kernel_launch_count()
        break
print('bonds %d..%d of L=%d chi=%s: %.2f ms, %d launches, E=%.12f' % (
    first, first + args.bonds - 1, L, psi.chi[first:first + args.bonds], ev0.elapsed_time(ev1), n0,
    eng.update_stats['E_total'][-1]))
