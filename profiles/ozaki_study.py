#!/usr/bin/env python
"""Numerical study (CPU, numpy) for the round-2 FP64-via-INT8 GEMM on the tcgen05 tensor cores (Ozaki scheme I):
how many int8 slices does a DMRG matvec GEMM need to reach the accuracy of the DMMA path?

    C = A B,   A = diag(2^ea) sum_p 2^(-beta p) A_p,   B = sum_q 2^(-beta q) B_q diag(2^eb),   A_p, B_q int8 (|.| < 2^beta)
    C ~= diag(2^ea) [ sum_{p+q < s} 2^(-beta (p+q)) (A_p B_q) ] diag(2^eb)        (A_p B_q exact in int32 for k 2^(2 beta) < 2^31)

ea / eb: exponents of the row maxima of A / column maxima of B.  s (s+1) / 2 integer GEMMs per FP64 GEMM.
Reported: max |C - C_ref| / (|A| |B|)_ij (the componentwise bound standard GEMM error analysis uses) and relative to
max|C_ref|, against float64 numpy (accumulation in long double for the reference).
"""
import sys

import numpy as np

BETA = 6          # magnitude bits per slice (int8 holds 7; 6 leaves headroom for a signed-digit representation)


def split(X, axis, s):
    """scale along `axis` (0: rows of A, 1: columns of B) and cut into `s` slices of BETA bits (truncation toward zero)"""
    mx = np.max(np.abs(X), axis=1 - axis, keepdims=True)
    e = np.where(mx > 0, np.ceil(np.log2(np.where(mx > 0, mx, 1.))), 0.)
    R = X / 2.**e                       # |R| <= 1
    slices = []
    for _ in range(s):
        R = R * 2.**BETA
        S = np.trunc(R)
        slices.append(S.astype(np.int64))
        R = R - S
    return e, slices


def ozaki_gemm(A, B, s):
    ea, As = split(A, 0, s)
    eb, Bs = split(B, 1, s)
    acc = np.zeros((A.shape[0], B.shape[1]), dtype=np.float64)
    n_gemm = 0
    for g in range(s - 1, -1, -1):      # p + q = g, small terms first
        part = np.zeros(acc.shape, dtype=np.int64)
        for p in range(g + 1):
            part += As[p] @ Bs[g - p]   # exact integer GEMM (int32 on the tensor cores: k 2^(2 BETA) < 2^31)
            n_gemm += 1
        acc += part.astype(np.float64) * 2.**(-BETA * (g + 2))
    return acc * 2.**ea * 2.**eb, n_gemm


def cases(rng, n):
    q1, _ = np.linalg.qr(rng.standard_normal((n, n)))
    q2, _ = np.linalg.qr(rng.standard_normal((n, n)))
    yield 'gaussian', rng.standard_normal((n, n)), rng.standard_normal((n, n))
    s = np.exp(-np.arange(n) / (n / 40.))
    yield 'theta-like (U S V^T, S over 17 decades)', (q1 * s) @ q2, rng.standard_normal((n, n))
    yield 'row-graded LHeff-like', rng.standard_normal((n, n)) * np.logspace(0, -8, n)[:, None], (q1 * s) @ q2


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    rng = np.random.default_rng(0)
    print('| case | slices s | int8 GEMMs | max err / (|A||B|) | max err / max|C| | float64 GEMM err / (|A||B|) |')
    print('|---|---:|---:|---:|---:|---:|')
    for name, A, B in cases(rng, n):
        ref = (A.astype(np.longdouble) @ B.astype(np.longdouble))
        bound = np.abs(A) @ np.abs(B)
        f64 = float(np.max(np.abs(A @ B - ref) / bound))
        for s in (6, 7, 8, 9, 10):
            C, ng = ozaki_gemm(A, B, s)
            err = np.abs(C - ref)
            print('| %s | %d | %d | %.1e | %.1e | %.1e |' % (name, s, ng, float(np.max(err / bound)),
                                                          float(np.max(err) / np.max(np.abs(ref))), f64))


if __name__ == '__main__':
    main()
