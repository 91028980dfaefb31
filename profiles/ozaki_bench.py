#!/usr/bin/env python
"""Timing of the int8 tensor-core FP64 product (csrc/ozaki.cu) on the two large GEMM shapes of the chi=1024 matvec,
next to the DMMA grouped GEMM on the same operands: split passes and multiply timed separately (CUDA events).

    python profiles/ozaki_bench.py [chi=1024]
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tenpy_b200 import backend


def timeit(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    ev0.record()
    for _ in range(reps):
        fn()
    ev1.record()
    torch.cuda.synchronize()
    return ev0.elapsed_time(ev1) / reps


def main():
    chi = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    lib = backend.get_lib()
    dev = lib.device
    d, Dm1 = 2, 2
    shapes = [('LP_rest.theta', Dm1 * chi, d * d * chi, chi), ('t2.RP_rest', d * d * chi, chi, Dm1 * chi),
              ('square', 2 * chi, 2 * chi, 2 * chi)]
    for name, m, n, k in shapes:
        g = torch.Generator(device=dev)
        g.manual_seed(1)
        A = torch.randn(m * k, dtype=torch.float64, device=dev, generator=g)
        B = torch.randn(k * n, dtype=torch.float64, device=dev, generator=g)
        C = torch.empty(m * n, dtype=torch.float64, device=dev)
        Cd = torch.empty(m * n, dtype=torch.float64, device=dev)
        flops = 2. * m * n * k
        # DMMA grouped GEMM on the same product
        one = np.array([0], dtype=np.int64)
        t_dmma = timeit(lambda: lib.grouped_gemm([m], [n], one, [0, 1], [k], one, one, A, B, Cd), reps=5)
        ref = Cd.clone()
        row = {'shape': name, 'm': m, 'n': n, 'k': k, 'dmma_ms': t_dmma, 'dmma_tflops': flops / t_dmma / 1e9}
        for s in (7, 8, 9):
            a_s = lib.ozaki_split(m, k, A, k, 1, s)
            b_s = lib.ozaki_split(n, k, B, 1, n, s)
            lib.ozaki_mm(m, n, k, s, a_s, b_s, C, n)
            lib.ozaki_check_abort()
            err = float((C - ref).abs().max() / ref.abs().max())
            t_sa = timeit(lambda: lib.ozaki_split(m, k, A, k, 1, s))
            t_sb = timeit(lambda: lib.ozaki_split(n, k, B, 1, n, s))
            t_mm = timeit(lambda: lib.ozaki_mm(m, n, k, s, a_s, b_s, C, n))
            npairs = s * (s + 1) // 2
            row['s%d' % s] = {'split_A_ms': t_sa, 'split_B_ms': t_sb, 'mm_ms': t_mm, 'mm_fp64_equiv_tflops': flops / t_mm / 1e9,
                              'int8_Tops': flops * npairs / t_mm / 1e9, 'int8_frac_of_4500': flops * npairs / t_mm / 1e9 / 4500.,
                              'total_fp64_equiv_tflops': flops / (t_mm + t_sa + t_sb) / 1e9,
                              'max_abs_diff_vs_dmma_rel': err}
        print(json.dumps(row))
        sys.stdout.flush()


if __name__ == '__main__':
    main()
