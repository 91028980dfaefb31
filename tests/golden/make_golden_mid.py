#!/usr/bin/env python
"""Mid-size golden DMRG runs (block-sparse path) generated with the UNMODIFIED reference (build container only):

    TENPY_NO_CYTHON=1 PYTHONPATH=/root/reference python tests/golden/make_golden_mid.py

Scaled-down versions of BASELINE.json configs[2] (XXZ / Sz) and configs[3] (Fermi-Hubbard / N,Sz)."""
import os
import sys
import time
import warnings

import numpy as np

os.environ.setdefault('TENPY_NO_CYTHON', '1')
sys.path.insert(0, os.environ.get('TENPY_REFERENCE', '/root/reference'))
warnings.simplefilter('ignore')
from tenpy.algorithms import dmrg  # noqa: E402
from tenpy.networks.mps import MPS  # noqa: E402
from tenpy.models.spins import SpinChain  # noqa: E402
from tenpy.models.hubbard import FermiHubbardChain  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
out = {}
params = {'mixer': True, 'mixer_params': {'amplitude': 1e-5, 'decay': 2., 'disable_after': 8}, 'max_E_err': 1e-11,
          'max_S_err': 1e-10, 'combine': True, 'max_sweeps': 40,
          'lanczos_params': {'P_tol': 1e-22, 'N_max': 40}}
t0 = time.time()
L = 32
M = SpinChain(dict(L=L, S=0.5, Jx=1., Jy=1., Jz=1.5, hz=0., bc_MPS='finite', conserve='Sz'))
psi = MPS.from_product_state(M.lat.mps_sites(), ['up', 'down'] * (L // 2), bc='finite')
eng = dmrg.TwoSiteDMRGEngine(psi, M, dict(params, trunc_params={'chi_max': 96, 'svd_min': 1e-10}))
E, _ = eng.run()
out['xxz32_E'] = np.float64(E)
out['xxz32_S'] = psi.entanglement_entropy()
out['xxz32_chi'] = np.array(psi.chi)
out['xxz32_sv_mid'] = np.sort(psi.get_SL(L // 2))[::-1]
print('xxz32', E, max(psi.chi), time.time() - t0)
L = 12
M = FermiHubbardChain(dict(L=L, t=1., U=4., mu=0., bc_MPS='finite', cons_N='N', cons_Sz='Sz'))
psi = MPS.from_product_state(M.lat.mps_sites(), ['up', 'down'] * (L // 2), bc='finite')
eng = dmrg.TwoSiteDMRGEngine(psi, M, dict(params, trunc_params={'chi_max': 1000, 'svd_min': 1e-5}))
E, _ = eng.run()
out['hub12_E'] = np.float64(E)
out['hub12_S'] = psi.entanglement_entropy()
out['hub12_chi'] = np.array(psi.chi)
out['hub12_sv_mid'] = np.sort(psi.get_SL(L // 2))[::-1]
print('hub12', E, max(psi.chi), time.time() - t0)
np.savez_compressed(os.path.join(HERE, 'dmrg_mid.npz'), **out)
