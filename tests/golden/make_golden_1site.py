#!/usr/bin/env python
"""Golden vectors of the single-site DMRG from the UNMODIFIED reference (build container only):

    TENPY_NO_CYTHON=1 PYTHONPATH=/root/reference python tests/golden/make_golden_1site.py

Cases: SingleSiteDMRGEngine with the DensityMatrixMixer (two-site mixing step) from product states -- TFI without
charges, XXZ with Sz, combine True / False -- and a mixer-free single-site refinement of a two-site DMRG state.
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402

from tenpy.algorithms import dmrg  # noqa: E402
from tenpy.algorithms.mps_common import OneSiteH  # noqa: E402
from tenpy.networks.mps import MPS  # noqa: E402
from tenpy.models.tf_ising import TFIChain  # noqa: E402
from tenpy.models.spins import SpinChain  # noqa: E402

warnings.simplefilter('ignore')


def main():
    out = {}
    # (1) TFI, no charges, single-site with DensityMatrixMixer from a product state
    for combine in (True, False):
        L = 12
        M = TFIChain(dict(L=L, J=1., g=1.1, bc_MPS='finite', conserve=None))
        psi = MPS.from_product_state(M.lat.mps_sites(), ['up'] * L, bc='finite')
        params = {'mixer': 'DensityMatrixMixer', 'mixer_params': {'amplitude': 1e-3, 'decay': 2., 'disable_after': 8},
                  'max_E_err': 1e-11, 'max_S_err': 1e-8, 'trunc_params': {'chi_max': 24, 'svd_min': 1e-10},
                  'combine': combine, 'max_sweeps': 24}
        eng = dmrg.SingleSiteDMRGEngine(psi, M, params)
        E, _ = eng.run()
        key = 'tfi_c%d' % int(combine)
        out[key + '_E'] = np.float64(E)
        out[key + '_S'] = psi.entanglement_entropy()
        out[key + '_chi'] = np.array(psi.chi)
        out[key + '_sweeps'] = np.int64(eng.sweeps)
        print(key, E, eng.sweeps, psi.chi)
    # (2) XXZ with Sz conservation
    L = 10
    M = SpinChain(dict(L=L, S=0.5, Jx=1., Jy=1., Jz=0.8, bc_MPS='finite', conserve='Sz'))
    psi = MPS.from_product_state(M.lat.mps_sites(), ['up', 'down'] * (L // 2), bc='finite')
    params = {'mixer': 'DensityMatrixMixer', 'mixer_params': {'amplitude': 1e-3, 'decay': 2., 'disable_after': 8},
              'max_E_err': 1e-11, 'max_S_err': 1e-8, 'trunc_params': {'chi_max': 32, 'svd_min': 1e-10},
              'combine': True, 'max_sweeps': 24}
    eng = dmrg.SingleSiteDMRGEngine(psi, M, params)
    E, _ = eng.run()
    out['xxz_E'] = np.float64(E)
    out['xxz_S'] = psi.entanglement_entropy()
    out['xxz_chi'] = np.array(psi.chi)
    print('xxz', E, eng.sweeps, psi.chi)
    # hot-path golden: OneSiteH.matvec in both directions on the converged state
    for move_right in (True, False):
        i0 = L // 2
        H = OneSiteH(eng.env, i0, combine=True, move_right=move_right)
        theta = H.combine_theta(psi.get_theta(i0, 1))
        tag = 'xxz_r' if move_right else 'xxz_l'
        mg.dump_array(tag + '_theta', theta, out)
        mg.dump_array(tag + '_Htheta', H.matvec(theta), out)
        if move_right:
            mg.dump_array(tag + '_LHeff', H.LHeff, out)
            mg.dump_array(tag + '_RP', H.RP, out)
        else:
            mg.dump_array(tag + '_RHeff', H.RHeff, out)
            mg.dump_array(tag + '_LP', H.LP, out)
    # (3) mixer-free single-site sweeps refining a truncated two-site state (chi fixed)
    L = 14
    M = TFIChain(dict(L=L, J=1., g=0.9, bc_MPS='finite', conserve=None))
    psi = MPS.from_product_state(M.lat.mps_sites(), ['up'] * L, bc='finite')
    eng2 = dmrg.TwoSiteDMRGEngine(psi, M, {'mixer': None, 'trunc_params': {'chi_max': 8, 'svd_min': 1e-12},
                                           'max_sweeps': 2, 'min_sweeps': 2, 'combine': True})
    eng2.run()
    out['ref_E2'] = np.float64(eng2.sweep_stats['E'][-1])
    eng1 = dmrg.SingleSiteDMRGEngine(psi, M, {'mixer': None, 'trunc_params': {'chi_max': 8, 'svd_min': 1e-12},
                                              'max_E_err': 1e-12, 'max_sweeps': 12, 'combine': True})
    E, _ = eng1.run()
    out['ref_E1'] = np.float64(E)
    out['ref_S1'] = psi.entanglement_entropy()
    print('refine', out['ref_E2'], E)
    # (4) SubspaceExpansion mixer (the reference's default for single-site DMRG): single-site TFI / XXZ-Sz from product
    #     states, and the two-site engine using it through Mixer.mix_and_decompose_2site
    mp = {'amplitude': 1e-3, 'decay': 2., 'disable_after': 8}
    L = 12
    M = TFIChain(dict(L=L, J=1., g=1.1, bc_MPS='finite', conserve=None))
    psi = MPS.from_product_state(M.lat.mps_sites(), ['up'] * L, bc='finite')
    eng = dmrg.SingleSiteDMRGEngine(psi, M, {'mixer': True, 'mixer_params': dict(mp), 'max_E_err': 1e-11,
                                             'max_S_err': 1e-8, 'trunc_params': {'chi_max': 24, 'svd_min': 1e-10},
                                             'combine': True, 'max_sweeps': 24})
    assert type(eng.mixer).__name__ == 'SubspaceExpansion' or eng.mixer is None
    E, _ = eng.run()
    out['se_tfi_E'], out['se_tfi_S'], out['se_tfi_chi'] = np.float64(E), psi.entanglement_entropy(), np.array(psi.chi)
    print('se_tfi', E, eng.sweeps, psi.chi)
    L = 10
    M = SpinChain(dict(L=L, S=0.5, Jx=1., Jy=1., Jz=0.8, bc_MPS='finite', conserve='Sz'))
    for key, Engine, combine in (('se_xxz1', dmrg.SingleSiteDMRGEngine, True), ('se_xxz1n', dmrg.SingleSiteDMRGEngine, False),
                                 ('se_xxz2', dmrg.TwoSiteDMRGEngine, True)):
        psi = MPS.from_product_state(M.lat.mps_sites(), ['up', 'down'] * (L // 2), bc='finite')
        eng = Engine(psi, M, {'mixer': 'SubspaceExpansion', 'mixer_params': dict(mp), 'max_E_err': 1e-11,
                              'max_S_err': 1e-8, 'trunc_params': {'chi_max': 32, 'svd_min': 1e-10},
                              'combine': combine, 'max_sweeps': 24})
        E, _ = eng.run()
        out[key + '_E'], out[key + '_S'], out[key + '_chi'] = np.float64(E), psi.entanglement_entropy(), np.array(psi.chi)
        print(key, E, eng.sweeps, psi.chi)
    np.savez_compressed(os.path.join(HERE, 'dmrg_1site.npz'), **out)


if __name__ == '__main__':
    main()
