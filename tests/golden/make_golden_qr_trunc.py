#!/usr/bin/env python
"""Golden vectors of the QR based truncation (reference tenpy/linalg/truncation.py:370-713) from the
UNMODIFIED reference (build container only):

    TENPY_NO_CYTHON=1 PYTHONPATH=/root/reference python tests/golden/make_golden_qr_trunc.py

Inputs: two-site wave functions of (not fully converged) DMRG states, perturbed so that the bond wants to grow;
outputs: S, truncation error, renormalisation and the reconstructed theta (gauge invariant) for move_right in {True,
False} x use_eig_based_svd in {False, True}.
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402

from tenpy.algorithms import dmrg  # noqa: E402
from tenpy.networks.mps import MPS  # noqa: E402
from tenpy.models.tf_ising import TFIChain  # noqa: E402
from tenpy.models.spins import SpinChain  # noqa: E402
from tenpy.linalg.truncation import decompose_theta_qr_based  # noqa: E402

npc = mg.npc
warnings.simplefilter('ignore')


def main():
    out = {}
    rng = np.random.default_rng(99)
    cases = []
    L = 10
    M = SpinChain(dict(L=L, S=0.5, Jx=1., Jy=1., Jz=0.8, bc_MPS='finite', conserve='Sz'))
    psi = MPS.from_product_state(M.lat.mps_sites(), ['up', 'down'] * (L // 2), bc='finite')
    dmrg.run(psi, M, dict(mixer=True, trunc_params=dict(chi_max=8, svd_min=1e-10), max_sweeps=3, min_sweeps=3))
    cases.append(('xxz', psi, M))
    M = TFIChain(dict(L=L, J=1., g=1.1, bc_MPS='finite', conserve=None))
    psi = MPS.from_product_state(M.lat.mps_sites(), ['up'] * L, bc='finite')
    dmrg.run(psi, M, dict(mixer=None, trunc_params=dict(chi_max=6, svd_min=1e-10), max_sweeps=3, min_sweeps=3))
    cases.append(('tfi', psi, M))
    n = 0
    for name, psi, M in cases:
        i0 = L // 2 - 1
        theta = psi.get_theta(i0, 2)
        # apply the two-site gate exp(-0.1 h) of the bond -> theta leaves the old bond space (what TEBD does)
        from tenpy.linalg import np_conserved as rnpc
        H2 = M.H_bond[i0 + 1]
        H2m = H2.combine_legs([['p0', 'p1'], ['p0*', 'p1*']], qconj=[+1, -1])
        Ub = rnpc.expm(-0.1 * H2m).split_legs()
        theta = rnpc.tensordot(Ub, theta, axes=[['p0*', 'p1*'], ['p0', 'p1']])
        theta = theta.combine_legs([['vL', 'p0'], ['p1', 'vR']], qconj=[+1, -1])
        theta.itranspose(['(vL.p0)', '(p1.vR)'])
        old_L, old_R = psi.get_B(i0, 'B'), psi.get_B(i0 + 1, 'B')
        for move_right in (True, False):
            for eig in (False, True):
                key = 'q%d' % n
                out[key + '_name'] = np.array(name)
                out[key + '_move_right'] = np.int64(move_right)
                out[key + '_eig'] = np.int64(eig)
                mg.dump_array(key + '_theta', theta, out)
                mg.dump_leg(key + '_oldleg', old_R.get_leg('vL'), out)
                out[key + '_qL'] = np.asarray(old_L.qtotal, dtype=np.int64)
                out[key + '_qR'] = np.asarray(old_R.qtotal, dtype=np.int64)
                tp = dict(chi_max=12, svd_min=1e-10)
                T_L, S, T_R, form, err, renorm = decompose_theta_qr_based(
                    old_L.qtotal, old_R.qtotal, old_R.get_leg('vL'), theta, move_right, 0.5, 1, eig, tp, True, True)
                out[key + '_S'] = np.asarray(S)
                out[key + '_eps'] = np.float64(err.eps)
                out[key + '_renorm'] = np.float64(renorm)
                out[key + '_form'] = np.array(form)
                if eig:
                    approx = npc.tensordot(T_L, T_R, ['vR', 'vL'])
                else:
                    approx = npc.tensordot(T_L.scale_axis(S, 'vR'), T_R, ['vR', 'vL'])
                mg.dump_array(key + '_approx', approx, out)
                print(key, name, move_right, eig, len(S), err.eps, renorm, form)
                n += 1
    out['n'] = np.int64(n)
    np.savez_compressed(os.path.join(HERE, 'qr_trunc.npz'), **out)


if __name__ == '__main__':
    main()
