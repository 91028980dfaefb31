#!/usr/bin/env python
"""Numerical study (CPU, numpy) of the outer iteration of the block-Jacobi SVD of csrc/svd.cu: how many sweeps does a
generic full-rank block need as a function of (a) how well the 32x32 pivot problem is diagonalised per round (inner sweeps
of the parallel-order Jacobi; 4 in the kernel today) and (b) QR preconditioning (Drmac-Veselic: iterate on the triangular
factor).  Emulates the kernel's structure: rows in blocks of 16, round-robin tournament over the block pairs, G = P P^T,
rotation of the 32-row panel, rows re-sorted by norm every sweep (de Rijk), convergence when a full sweep touched nothing
(scaled off-diagonal of every pair below 1e-14 sqrt(p), the library's threshold).

    python profiles/jacobi_sweeps_study.py [n=256]
"""
import sys

import numpy as np

JB = 16


def inner_eig(G, sweeps):
    """parallel-order cyclic Jacobi on the symmetric 32x32 G, `sweeps` sweeps (None: exact eigh); returns Q (G ~ Q L Q^T)"""
    n = G.shape[0]
    if sweeps is None:
        w, Q = np.linalg.eigh(G)
        return Q[:, ::-1]
    G = G.copy()
    Q = np.eye(n)
    for _ in range(sweeps):
        rotated = False
        for step in range(n - 1):
            J = np.eye(n)
            for t in range(n // 2):
                a, b = (n - 1, step) if t == 0 else ((step + t) % (n - 1), (step - t + n - 1) % (n - 1))
                p, q = min(a, b), max(a, b)
                gpp, gqq, gpq = G[p, p], G[q, q], G[p, q]
                if abs(gpq) > 1e-15 * np.sqrt(abs(gpp * gqq)):
                    aa, bb = gqq - gpp, 2. * gpq
                    hh = np.hypot(aa, bb)
                    tt = bb / (aa + hh) if aa >= 0 else bb / (aa - hh)
                    c = 1. / np.sqrt(1. + tt * tt)
                    s = tt * c
                    J[p, p], J[q, q], J[p, q], J[q, p] = c, c, s, -s
                    rotated = True
            G = J.T @ G @ J
            Q = Q @ J
        if not rotated:
            break
    order = np.argsort(-np.diag(G))
    return Q[:, order]


def block_jacobi_sweeps(Y, inner, max_sweeps=40):
    """one-sided block Jacobi on the rows of Y; returns (sweeps, singular values)"""
    Y = Y.copy()
    q, p = Y.shape
    nb = q // JB
    tol = 1e-14 * np.sqrt(p)
    for sweep in range(1, max_sweeps + 1):
        Y = Y[np.argsort(-np.linalg.norm(Y, axis=1))]          # de Rijk
        touched = 0
        for r in range(nb - 1):
            for j in range(nb // 2):
                a, b = (nb - 1, r) if j == 0 else ((r + j) % (nb - 1), (r - j + nb - 1) % (nb - 1))
                rows = np.r_[min(a, b) * JB:(min(a, b) + 1) * JB, max(a, b) * JB:(max(a, b) + 1) * JB]
                P = Y[rows]
                G = P @ P.T
                d = np.sqrt(np.outer(np.diag(G), np.diag(G)))
                off = np.max(np.abs(G - np.diag(np.diag(G))) / np.where(d > 0, d, 1.))
                if off > tol:
                    Y[rows] = inner_eig(G, inner).T @ P
                    touched += 1
        if touched == 0:
            return sweep, np.sort(np.linalg.norm(Y, axis=1))[::-1]
    return max_sweeps, np.sort(np.linalg.norm(Y, axis=1))[::-1]


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    rng = np.random.default_rng(1)
    A = rng.standard_normal((n, n))
    Sref = np.linalg.svd(A, compute_uv=False)
    print('| start | pivot problem per round | outer sweeps | max |S - S_lapack| / S_max |')
    print('|---|---|---:|---:|')
    Qf, R = np.linalg.qr(A)
    Qp, Rp, piv = __import__('scipy.linalg', fromlist=['qr']).qr(A, pivoting=True)
    for start, Y in (('A (rows)', A), ('R^T of A = Q R', R.T), ('R^T of A P = Q R (column pivoting)', Rp.T), ('R of A = Q R', R)):
        for inner in (2, 4, 7, None):
            sw, S = block_jacobi_sweeps(Y, inner)
            print('| %s | %s | %d | %.1e |' % (start, 'exact' if inner is None else '%d inner sweeps' % inner, sw,
                                             float(np.max(np.abs(S - Sref)) / Sref[0])))


if __name__ == '__main__':
    main()
