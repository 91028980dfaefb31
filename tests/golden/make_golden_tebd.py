#!/usr/bin/env python
"""Golden TEBD vectors generated with the UNMODIFIED reference (build container only):

    TENPY_NO_CYTHON=1 PYTHONPATH=/root/reference python tests/golden/make_golden_tebd.py

A fixed number of imaginary-time steps (no convergence loop, so the amount of work is deterministic) through
`TEBDEngine.update_imag` (sweeps) and `TEBDEngine.evolve` (brick wall, orders 1/2/4); bond energies, entropies,
bond dimensions and the H_bond / U_bond operators themselves are stored."""
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
from tenpy.models.hubbard import FermiHubbardChain  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
out = {}


def record(tag, M, psi, eng):
    out[tag + '_Ebond'] = np.asarray(M.bond_energies(psi), dtype=np.float64)
    out[tag + '_S'] = psi.entanglement_entropy()
    out[tag + '_chi'] = np.array(psi.chi)
    out[tag + '_norm'] = np.float64(psi.norm)
    out[tag + '_eps'] = np.float64(eng.trunc_err.eps)
    print(tag, float(np.sum(out[tag + '_Ebond'])), max(psi.chi), psi.norm)


def cases():
    L = 10
    yield 'tfi', TFIChain(dict(L=L, J=1., g=1.2, bc_MPS='finite', conserve=None)), ['up'] * L
    yield 'tfip', TFIChain(dict(L=L, J=1., g=0.8, bc_MPS='finite', conserve='parity')), ['up'] * L
    yield 'xxz', SpinChain(dict(L=L, S=0.5, Jx=1., Jy=1., Jz=1.3, hz=0., bc_MPS='finite', conserve='Sz')), \
        ['up', 'down'] * (L // 2)
    L = 6
    yield 'hub', FermiHubbardChain(dict(L=L, t=1., U=3., mu=0., bc_MPS='finite', cons_N='N', cons_Sz='Sz')), \
        ['up', 'down'] * (L // 2)


for name, M, state in cases():
    sites = M.lat.mps_sites()
    L = len(sites)
    hb = M.H_bond[L // 2].to_ndarray()                    # p0, p0*, p1, p1*
    out[name + '_Hbond_mid'] = hb
    out[name + '_Hbond_first'] = M.H_bond[1].to_ndarray()
    # 1) imaginary-time sweeps (what run_GS does for finite chains at second order)
    psi = MPS.from_product_state(sites, state, bc='finite')
    eng = tebd.TEBDEngine(psi, M, {'trunc_params': {'chi_max': 32, 'svd_min': 1e-8}})
    eng.calc_U(2, 0.05, type_evo='imag')
    out[name + '_U_half_mid'] = eng._U[0][L // 2].to_ndarray()   # p0, p1, p0*, p1*
    eng.update_imag(30, call_canonical_form=False)
    record(name + '_imag', M, psi, eng)
    # 2) brick-wall evolution, orders 1, 2, 4 (imaginary time: state leaves the canonical form slightly; the
    #    numbers are still a deterministic function of the update rule)
    for order in (1, 2, 4):
        psi = MPS.from_product_state(sites, state, bc='finite')
        eng = tebd.TEBDEngine(psi, M, {'trunc_params': {'chi_max': 32, 'svd_min': 1e-8}})
        eng.calc_U(order, 0.02, type_evo='imag')
        eng.evolve(6, 0.02)
        record('{0}_o{1}'.format(name, order), M, psi, eng)

np.savez_compressed(os.path.join(HERE, 'tebd.npz'), **out)
