"""dev check (NOT collected by pytest; run by hand on the GPU box first thing in round 2):

    python tests/dev_eig_v2_gpu_check.py

Switches the pivot eigen-solver of the Jacobi rounds to `jacobi_eig_kernel_v2` (b200_svd_set_eig_variant(2)), compares
SVD / eigh results with version 1 and with numpy on a set of shapes, and times a generic 2048x2048 block with both."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, '.')
from tenpy_b200 import backend
from tenpy_b200.linalg import np_conserved as npc


def main():
    lib = backend.get_lib()
    rng = np.random.default_rng(3)
    ok = True
    for shape in [(5, 5), (33, 47), (64, 64), (200, 150), (512, 512)]:
        A = rng.standard_normal(shape) * np.logspace(0, -6, shape[1])[None, :]
        a = npc.Array.from_ndarray_trivial(A)
        res = {}
        for variant in (1, 2):
            old = lib.svd_set_eig_variant(variant)
            try:
                U, S, VH = npc.svd(a)
            finally:
                lib.svd_set_eig_variant(old)
            rec = np.max(np.abs(U.to_ndarray() @ np.diag(S) @ VH.to_ndarray() - A))
            sd = np.max(np.abs(np.sort(S)[::-1] - np.linalg.svd(A, compute_uv=False)))
            orth = np.max(np.abs(U.to_ndarray().T @ U.to_ndarray() - np.eye(min(shape))))
            res[variant] = (rec, sd, orth, npc.svd_stats['jacobi_sweeps'][-1])
        print(shape, 'v1 rec %.1e dS %.1e orth %.1e sweeps %d | v2 rec %.1e dS %.1e orth %.1e sweeps %d' % (res[1] + res[2]))
        ok = ok and res[2][0] < 1e-11 and res[2][1] < 1e-11 and res[2][2] < 1e-11
    n = 2048
    th = npc.Array.from_ndarray_trivial(rng.standard_normal((n, n)))
    for variant, inner in ((1, 4), (2, 4), (2, 3), (2, 2), (2, 1)):
        old = lib.svd_set_eig_variant(variant)
        old_in = lib.svd_set_eig_inner_sweeps(inner)
        try:
            npc.svd(th)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            U, S, VH = npc.svd(th)
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) * 1e3
            rec = float(npc.norm(npc.tensordot(U.scale_axis(S, 1), VH, axes=1) - th) / npc.norm(th))
            print('2048x2048 generic, eig variant %d, %d inner sweeps: %.1f ms, %d outer sweeps, rec.err %.1e' % (
                variant, inner, dt, npc.svd_stats['jacobi_sweeps'][-1], rec))
        finally:
            lib.svd_set_eig_variant(old)
            lib.svd_set_eig_inner_sweeps(old_in)
    # Householder block QR (b200_block_qr_f64) against the Gram-Schmidt route and numpy
    for shape in [(7, 4), (4, 7), (64, 64), (300, 130), (512, 512)]:
        A = rng.standard_normal(shape)
        a = npc.Array.from_ndarray_trivial(A)
        res = {}
        for method in ('cgs2', 'householder'):
            old_m = npc.qr_method
            npc.qr_method = method
            try:
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                Q, R = npc.qr(a)
                torch.cuda.synchronize()
                dt = (time.perf_counter() - t0) * 1e3
            finally:
                npc.qr_method = old_m
            q, r = Q.to_ndarray(), R.to_ndarray()
            res[method] = (np.max(np.abs(q @ r - A)), np.max(np.abs(q.T @ q - np.eye(q.shape[1]))), dt, q, r)
        dq = np.max(np.abs(res['cgs2'][3] - res['householder'][3]))
        print(shape, 'qr cgs2 rec %.1e orth %.1e %.1f ms | householder rec %.1e orth %.1e %.1f ms | dQ %.1e' % (
            res['cgs2'][:3] + res['householder'][:3] + (dq,)))
        ok = ok and res['householder'][0] < 1e-11 and res['householder'][1] < 1e-12 and dq < 1e-9
    print('ok' if ok else 'FAILED')
    return 0 if ok else 1


if __name__ == '__main__':
    sys.exit(main())
