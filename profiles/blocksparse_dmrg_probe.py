#!/usr/bin/env python
"""BASELINE.json configs[2] / [3] end to end: two-site DMRG with charge conservation from a product state, bond dimension
ramped up to `--chi`, then `--timed` sweeps at the final bond dimension timed with CUDA events.

    python profiles/blocksparse_dmrg_probe.py xxz     --L 100 --chi 1024 --ramp 6 --timed 2     # SpinChain, U(1) Sz
    python profiles/blocksparse_dmrg_probe.py hubbard --L 64  --chi 2048 --ramp 6 --timed 2     # U(1) x U(1) (N, Sz)

Prints one JSON line: seconds per timed sweep, E, S_mid, max chi, sectors / blocks of the centre theta, Lanczos and SVD
statistics, kernel-family times of the last sweep.  (Not the benchmark of bench.py: the state is a real DMRG state, the
number of Lanczos iterations is the reference's adaptive one.)
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('model', choices=['xxz', 'hubbard'])
    ap.add_argument('--L', type=int, default=100)
    ap.add_argument('--chi', type=int, default=1024)
    ap.add_argument('--ramp', type=int, default=6, help='sweeps with growing chi (doubling from 32) and mixer')
    ap.add_argument('--timed', type=int, default=2)
    args = ap.parse_args()
    import torch
    from tenpy_b200 import backend
    from tenpy_b200.models import SpinChain, FermiHubbardChain
    from tenpy_b200.networks.mps import MPS
    from tenpy_b200.algorithms import dmrg
    from tenpy_b200.algorithms.mps_common import TwoSiteH
    lib = backend.get_lib()
    cuda = lib.device.type == 'cuda'
    L = args.L
    if args.model == 'xxz':
        M = SpinChain({'L': L, 'Jx': 1., 'Jy': 1., 'Jz': 1., 'conserve': 'Sz'})
    else:
        M = FermiHubbardChain({'L': L, 't': 1., 'U': 4., 'mu': 0.})
    psi = MPS.from_product_state(M.lat_sites, ['up', 'down'] * (L // 2))
    chis = [min(args.chi, 32 * 2**k) for k in range(args.ramp)]
    chis[-1] = args.chi
    opts = {'mixer': True, 'mixer_params': {'amplitude': 1e-4, 'decay': 2., 'disable_after': args.ramp},
            'combine': True, 'trunc_params': {'chi_max': chis[0], 'svd_min': 1e-12}}
    eng = dmrg.TwoSiteDMRGEngine(psi, M, opts)
    t_ramp = []
    for c in chis:
        eng.trunc_params['chi_max'] = c
        t0 = time.perf_counter()
        eng.sweep()
        lib.synchronize()
        t_ramp.append(round(time.perf_counter() - t0, 3))
    eng.mixer_deactivate()
    times = []
    fam = {}
    from tenpy_b200.linalg import np_conserved as npc_mod
    plans_before, wall_last = 0, None
    for k in range(args.timed):
        if k == args.timed - 1 and cuda:
            lib.profile = {}
            plans_before = len(npc_mod._PLAN_CACHE)
        if cuda:
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
        t0 = time.perf_counter()
        eng.sweep()
        if cuda:
            ev1.record()
        lib.synchronize()
        wall_last = time.perf_counter() - t0
        times.append(ev0.elapsed_time(ev1) / 1e3 if cuda else time.perf_counter() - t0)
    detail = {}
    if cuda and lib.profile is not None:
        fam = {k: round(v[1], 1) for k, v in lib.profile_summary().items()}
        det = lib.profile_detail()
        lib.profile = None
        # contraction efficiency by size class: where does the GEMM family time go?
        g = [(ms, info) for ms, info in det.get('gemm', []) if info]
        bins = [(0, 1e6), (1e6, 1e7), (1e7, 1e8), (1e8, 1e9), (1e9, 1e10), (1e10, 1e13)]
        hist = []
        for lo, hi in bins:
            sel = [(ms, i) for ms, i in g if lo <= i[0] < hi]
            if sel:
                tms, tfl = sum(x[0] for x in sel), sum(x[1][0] for x in sel)
                hist.append({'flop_range': [lo, hi], 'calls': len(sel), 'ms': round(tms, 2), 'gflop': round(tfl / 1e9, 2),
                             'tflops': round(tfl / tms / 1e9, 3) if tms else None,
                             'mean_pairs': round(float(np.mean([x[1][1] for x in sel])), 1)})
        slow = sorted(g, key=lambda x: -x[0])[:8]
        detail = {'gemm_by_flops': hist, 'gemm_slowest': [{'ms': round(ms, 3), 'flops': i[0], 'pairs': i[1], 'blocks_out': i[2]}
                                                          for ms, i in slow],
                  'calls': {k: len(v) for k, v in det.items()},
                  'wall_s_last_sweep': wall_last, 'plan_cache_size': len(npc_mod._PLAN_CACHE),
                  'plans_built_last_sweep': len(npc_mod._PLAN_CACHE) - plans_before}
    i0 = L // 2 - 1
    H = TwoSiteH(eng.env, i0, combine=True)
    theta = H.combine_theta(psi.get_theta(i0, 2))
    from tenpy_b200.linalg.np_conserved import svd_stats
    nb = 2 * (L - 2)
    print(json.dumps({'model': args.model, 'L': L, 'chi_max': args.chi, 'chi_reached': int(max(psi.chi)),
                      'ramp_chis': chis, 'ramp_sweep_s': t_ramp, 'timed_sweep_s': [round(t, 4) for t in times],
                      'E': float(eng.update_stats['E_total'][-1]), 'S_mid': float(psi.entanglement_entropy()[L // 2 - 1]),
                      'theta_blocks': int(theta.stored_blocks), 'theta_shape': list(theta.shape),
                      'theta_largest_block': [int(x) for x in theta._layout.shapes[np.argmax(theta._layout.sizes)]],
                      'bond_sectors': int(psi.get_B(L // 2).get_leg('vL').block_number),
                      'N_lanczos_mean': float(np.mean(eng.update_stats['N_lanczos'][-nb:])),
                      'svd_jacobi_sweeps_mean': float(np.mean(svd_stats['jacobi_sweeps'][-nb:])),
                      'family_ms_last_sweep': fam, 'detail': detail}))


if __name__ == '__main__':
    main()
