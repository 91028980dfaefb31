#!/usr/bin/env python
"""cProfile of the HOST side of one benchmark sweep (TFI L, chi; synthetic saturated MPS, 3 warm-up sweeps): which Python
functions the interpreter spends its time in while the GPU works asynchronously (tottime of a function that blocks on the
device includes the wait).

    python profiles/host_profile.py [L=100] [chi=1024]
"""
import cProfile
import io
import os
import pstats
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
    backend.get_lib()
    model = TFIChain({'L': L, 'J': 1., 'g': 1., 'conserve': None})
    psi = bench.synthetic_mps(model, L, chi, 2, seed=0)
    eng = dmrg.TwoSiteDMRGEngine(psi, model, {
        'mixer': None, 'combine': True, 'diag_method': 'lanczos', 'svd_warm_start': False,
        'trunc_params': {'chi_max': chi, 'svd_min': 1e-45, 'trunc_cut': None, 'svd_deflation_tol': 1e-10},
        'lanczos_params': {'N_min': 10, 'N_max': 10}})
    for _ in range(3):
        eng.sweep()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    eng.sweep()
    torch.cuda.synchronize()
    print('sweep without profiler: %.3f s' % (time.perf_counter() - t0))
    pr = cProfile.Profile()
    pr.enable()
    eng.sweep()
    torch.cuda.synchronize()
    pr.disable()
    for key in ('tottime', 'cumulative'):
        out = io.StringIO()
        pstats.Stats(pr, stream=out).sort_stats(key).print_stats(45)
        print(out.getvalue())


if __name__ == '__main__':
    main()
