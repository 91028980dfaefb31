#!/usr/bin/env python
"""How long does the GPU wait for the host BETWEEN the Jacobi sweeps of the block SVD (convergence test, active-set
re-ordering, descriptor updates) in a converged benchmark sweep?  Runs warm-up sweeps of the benchmark workload, then one
sweep with B200_JACOBI_DEBUG=1 and sums the `wait` / `host gap` figures the library prints per Jacobi sweep.

    python profiles/svd_gap_probe.py [L=40] [chi=1024] 2> log
"""
import json
import os
import re
import sys
import tempfile

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from tenpy_b200 import backend  # noqa: E402
from tenpy_b200.algorithms import dmrg  # noqa: E402
from tenpy_b200.models import TFIChain  # noqa: E402


def main():
    L = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    chi = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    backend.get_lib()
    model = TFIChain({'L': L, 'J': 1., 'g': 1., 'conserve': None})
    psi = bench.synthetic_mps(model, L, chi, 2, seed=0)
    eng = dmrg.TwoSiteDMRGEngine(psi, model, {
        'mixer': None, 'combine': True, 'diag_method': 'lanczos', 'svd_warm_start': False,
        'trunc_params': {'chi_max': chi, 'svd_min': 1e-45, 'trunc_cut': None, 'svd_deflation_tol': 1e-10},
        'lanczos_params': {'N_min': 10, 'N_max': 10}})
    for _ in range(4):
        eng.sweep()
    torch.cuda.synchronize()
    # redirect the C library's stderr into a file for one sweep
    tmp = tempfile.TemporaryFile(mode='w+b')
    sys.stderr.flush()
    saved = os.dup(2)
    os.dup2(tmp.fileno(), 2)
    os.environ['B200_JACOBI_DEBUG'] = '1'
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    eng.sweep()
    ev1.record()
    torch.cuda.synchronize()
    del os.environ['B200_JACOBI_DEBUG']
    os.dup2(saved, 2)
    tmp.seek(0)
    text = tmp.read().decode()
    waits = [float(x) for x in re.findall(r'wait ([0-9.]+) ms', text)]
    gaps = [float(x) for x in re.findall(r'host gap ([0-9.]+) ms', text)]
    svd = re.findall(r'\[svd\] prep\+init ([0-9.]+) ms, jacobi ([0-9.]+) ms \((\d+) sweeps\), sort\+finalize ([0-9.]+) ms', text)
    out = {'L': L, 'chi': chi, 'sweep_s': ev0.elapsed_time(ev1) / 1e3, 'svd_calls': len(svd), 'jacobi_sweeps': len(waits),
           'sum_wait_ms': sum(waits), 'sum_host_gap_ms': sum(gaps),
           'sum_prep_init_ms': sum(float(a) for a, _, _, _ in svd), 'sum_jacobi_ms': sum(float(b) for _, b, _, _ in svd),
           'sum_sort_finalize_ms': sum(float(d) for _, _, _, d in svd),
           'sample': text.splitlines()[len(text.splitlines()) // 2 - 6:len(text.splitlines()) // 2 + 6]}
    print(json.dumps(out))


if __name__ == '__main__':
    main()
