#!/usr/bin/env python
"""block SVD timings with both pivot eigen-solvers (csrc/svd.cu): generic dense blocks and the C3 / C4 shaped block sets.
    python profiles/svd_variants.py"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tenpy_b200 import backend


def run(lib, shapes, reps=3, decay=None):
    """`decay`: the first block gets singular values exp(-i / decay) (a numerically low-rank DMRG-like wave function)"""
    a_off, u_off, s_off, v_off = [], [], [], []
    ao = uo = so = vo = 0
    for m, n in shapes:
        k = min(m, n)
        a_off.append(ao), u_off.append(uo), s_off.append(so), v_off.append(vo)
        ao += m * n + (-(m * n)) % 16
        uo += m * k + (-(m * k)) % 16
        so += k + (-k) % 16
        vo += k * n + (-(k * n)) % 16
    g = torch.Generator(device=lib.device)
    g.manual_seed(5)
    A = torch.randn(ao, dtype=torch.float64, device=lib.device, generator=g)
    if decay:
        m, n = shapes[0]
        k = min(m, n)
        q1, _ = torch.linalg.qr(torch.randn(m, k, dtype=torch.float64, device=lib.device, generator=g))
        q2, _ = torch.linalg.qr(torch.randn(n, k, dtype=torch.float64, device=lib.device, generator=g))
        sv = torch.exp(-torch.arange(k, dtype=torch.float64, device=lib.device) / decay)
        A[:m * n] = ((q1 * sv) @ q2.T).reshape(-1)
    U, S, V = backend.zeros(uo), backend.zeros(so), backend.zeros(vo)
    ms, ns = [s[0] for s in shapes], [s[1] for s in shapes]
    info, _, _ = lib.block_svd(ms, ns, a_off, u_off, s_off, v_off, A, U, S, V)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    ev0.record()
    for _ in range(reps):
        lib.block_svd(ms, ns, a_off, u_off, s_off, v_off, A, U, S, V)
    ev1.record()
    torch.cuda.synchronize()
    m, n = shapes[0]
    k = min(m, n)
    A0, U0, S0, V0 = A[:m * n].view(m, n), U[:m * k].view(m, k), S[:k], V[:k * n].view(k, n)
    nz = int((S0 > 0).sum())
    rec = float(torch.linalg.norm(A0 - (U0 * S0) @ V0) / torch.linalg.norm(A0))
    orth = float(torch.linalg.norm(U0[:, :nz].T @ U0[:, :nz] - torch.eye(nz, dtype=torch.float64, device=lib.device)))
    sref = torch.linalg.svdvals(A0)
    serr = float((torch.sort(S0, descending=True)[0] - sref).abs().max() / sref[0])
    return ev0.elapsed_time(ev1) / reps, int(max(info)), (rec, orth, serr)


def main():
    lib = backend.get_lib()
    rng = np.random.default_rng(0)
    c3 = [(332, 177), (290, 291), (145, 145), (145, 144), (85, 85), (85, 84), (24, 24), (24, 25), (7, 8)]
    c4 = [(int(x), int(y)) for x, y in zip(rng.integers(3, 250, 39), rng.integers(3, 250, 39))]
    cases = [('generic 512^2', [(512, 512)]), ('generic 1024^2', [(1024, 1024)]), ('generic 2048^2', [(2048, 2048)]),
             ('C3-shaped set (9 blocks <= 332x177)', c3), ('C4-shaped set (39 blocks <= 250)', c4),
             ('XXZ chi=1024 set (9 blocks <= 718)', [(718, 718), (600, 601), (400, 400), (399, 400), (180, 181), (180, 180),
                                                     (50, 50), (49, 50), (8, 8)])]
    cases += [('low rank 2048^2 (sigma_i = exp(-i/30), deflation 1e-10)', [(2048, 2048)]),
              ('low rank 4 x 768^2 (sigma_i = exp(-i/20), deflation 1e-10)', [(768, 768)] * 4)]
    for name, shapes in cases:
        row = {'case': name}
        low = name.startswith('low rank')
        old_tol = lib.svd_set_deflation_tol(1e-10) if low else None
        for v, inner in ((3, 2), (3, 1), (3, 0)):
            old = lib.svd_set_eig_variant(v)
            old_in = lib.svd_set_eig_inner_sweeps(inner)
            ms, sweeps, acc = run(lib, shapes, reps=2 if shapes[0][0] >= 2048 else 3, decay=(30 if '2048' in name else 20) if low else None)
            lib.svd_set_eig_variant(old)
            lib.svd_set_eig_inner_sweeps(old_in)
            tag = 'v%d_%s' % (v, 'cross' if inner == 0 else 'in%d' % inner)
            row[tag + '_ms'] = round(ms, 3)
            row[tag + '_sweeps'] = sweeps
            row[tag + '_rec_orth_serr'] = ['%.1e' % x for x in acc]
        if low:
            lib.svd_set_deflation_tol(old_tol)
        print(json.dumps(row))
        sys.stdout.flush()


if __name__ == '__main__':
    main()
