#!/usr/bin/env python
"""Golden vectors of the QR based TEBD engine (reference tebd.py:619) from the UNMODIFIED reference (build container):

    TENPY_NO_CYTHON=1 PYTHONPATH=/root/reference python tests/golden/make_golden_tebd_qr.py

Fixed numbers of imaginary-time steps: `update_imag` sweeps (update_bond_imag) and brick-wall `evolve` (update_bond)."""
import os
import sys
import warnings

import numpy as np

os.environ.setdefault('TENPY_NO_CYTHON', '1')
sys.path.insert(0, os.environ.get('TENPY_REFERENCE', '/root/reference'))
warnings.simplefilter('ignore')
from tenpy.algorithms import tebd  # noqa: E402
from tenpy.networks.mps import MPS  # noqa: E402
from tenpy.models.tf_ising import TFIChain  # noqa: E402
from tenpy.models.spins import SpinChain  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
out = {}
OPTS = {'trunc_params': {'chi_max': 24, 'svd_min': 1e-8}, 'cbe_expand': 0.1, 'cbe_expand_0': 0.5,
        'cbe_min_block_increase': 2, 'compute_err': True}


def record(tag, M, psi, eng):
    out[tag + '_Ebond'] = np.asarray(M.bond_energies(psi), dtype=np.float64)
    out[tag + '_S'] = psi.entanglement_entropy()
    out[tag + '_chi'] = np.array(psi.chi)
    out[tag + '_norm'] = np.float64(psi.norm)
    out[tag + '_eps'] = np.float64(eng.trunc_err.eps)
    print(tag, float(np.sum(out[tag + '_Ebond'])), psi.chi, psi.norm, eng.trunc_err.eps)


L = 10
for name, M, state in (('tfi', TFIChain(dict(L=L, J=1., g=1.2, bc_MPS='finite', conserve=None)), ['up'] * L),
                       ('xxz', SpinChain(dict(L=L, S=0.5, Jx=1., Jy=1., Jz=1.3, hz=0., bc_MPS='finite', conserve='Sz')),
                        ['up', 'down'] * (L // 2))):
    sites = M.lat.mps_sites()
    psi = MPS.from_product_state(sites, state, bc='finite')
    eng = tebd.QRBasedTEBDEngine(psi, M, dict(OPTS))
    eng.calc_U(2, 0.05, type_evo='imag')
    eng.update_imag(20, call_canonical_form=False)
    record(name + '_imag', M, psi, eng)
    psi = MPS.from_product_state(sites, state, bc='finite')
    eng = tebd.QRBasedTEBDEngine(psi, M, dict(OPTS))
    eng.calc_U(2, 0.02, type_evo='imag')
    eng.evolve(6, 0.02)
    record(name + '_o2', M, psi, eng)
np.savez_compressed(os.path.join(HERE, 'tebd_qr.npz'), **out)
