#!/usr/bin/env python
"""The UNMODIFIED reference drivers (tenpy.algorithms.dmrg / tebd, tenpy.networks.*, tenpy.models.*) running on the
tenpy_b200 engine (tenpy_b200.dropin).  Executed in its own process by tests/test_dropin_engine.py because the seeding has
to happen before the first ``import tenpy``.

    python tests/dropin/run_reference_drivers.py fake|cuda [case ...]

Prints one JSON line per case.  The numbers are compared with the reference running on its own NumPy engine
(tests/golden/dropin.json, written by ``--golden`` with the plain reference).
"""
import json
import os
import sys
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
warnings.filterwarnings('ignore')


def setup(mode):
    if mode == 'golden':             # the plain reference, to write the expected numbers
        from tenpy_b200 import dropin
        sys.path.insert(0, dropin.reference_path())
        return None
    from tenpy_b200 import backend, dropin
    if mode == 'fake':
        from fake_device import FakeDeviceLib
        backend.use_library(FakeDeviceLib())
    else:
        from tenpy_b200._lib import DeviceLib
        backend.use_library(DeviceLib())
    path = dropin.install()
    assert path is not None, 'reference not found (baseline/_ref)'
    import tenpy
    import tenpy.linalg.np_conserved as npc
    assert npc.__name__ == 'tenpy_b200.linalg.np_conserved', npc.__name__
    from tenpy.algorithms import dmrg
    assert dmrg.npc is npc
    assert os.path.realpath(dmrg.__file__).startswith(os.path.realpath(path)), dmrg.__file__
    return dropin


def case_tfi_dmrg(dropin):
    """config 0: examples/d_dmrg.py TFIChain L=20 chi=50 two-site DMRG (E = -25.1077971116238)"""
    from tenpy.models.tf_ising import TFIChain
    from tenpy.networks.mps import MPS
    from tenpy.algorithms import dmrg
    M = TFIChain({'L': 20, 'J': 1., 'g': 1., 'bc_MPS': 'finite', 'conserve': None})
    psi = MPS.from_product_state(M.lat.mps_sites(), ['up'] * 20, bc='finite')
    info = dmrg.run(psi, M, {'mixer': None, 'max_E_err': 1.e-10, 'trunc_params': {'chi_max': 50, 'svd_min': 1.e-10},
                             'combine': True})
    return {'E': float(info['E']), 'S_mid': float(psi.entanglement_entropy()[9]), 'chi': [int(c) for c in psi.chi]}


def case_xxz_dmrg_mixer(dropin):
    """SpinChain L=16 with Sz conservation, density-matrix mixer, ragged charge blocks"""
    from tenpy.models.spins import SpinChain
    from tenpy.networks.mps import MPS
    from tenpy.algorithms import dmrg
    L = 16
    M = SpinChain({'L': L, 'S': 0.5, 'Jx': 1., 'Jy': 1., 'Jz': 1., 'bc_MPS': 'finite', 'conserve': 'Sz'})
    psi = MPS.from_product_state(M.lat.mps_sites(), ['up', 'down'] * (L // 2), bc='finite')
    info = dmrg.run(psi, M, {'mixer': True, 'mixer_params': {'amplitude': 1.e-5, 'decay': 2., 'disable_after': 6},
                             'max_E_err': 1.e-11, 'max_S_err': 1.e-8, 'max_sweeps': 20, 'combine': True,
                             'trunc_params': {'chi_max': 60, 'svd_min': 1.e-10}})
    return {'E': float(info['E']), 'S_mid': float(psi.entanglement_entropy()[L // 2 - 1]), 'chi': [int(c) for c in psi.chi]}


def case_tfi_dmrg_fast_engine(dropin):
    """the reference engine with the device-optimised effective Hamiltonian plugged in at `EffectiveH`"""
    from tenpy.models.tf_ising import TFIChain
    from tenpy.networks.mps import MPS
    M = TFIChain({'L': 20, 'J': 1., 'g': 1., 'bc_MPS': 'finite', 'conserve': None})
    psi = MPS.from_product_state(M.lat.mps_sites(), ['up'] * 20, bc='finite')
    if dropin is None:
        from tenpy.algorithms.dmrg import TwoSiteDMRGEngine as Engine
    else:
        Engine = dropin.fast_two_site_engine()
        Engine.EffectiveH.SPLIT_MIN_BLOCK = 1          # force the split / identity-environment route on small blocks
    eng = Engine(psi, M, {'mixer': None, 'max_E_err': 1.e-10, 'trunc_params': {'chi_max': 50, 'svd_min': 1.e-10},
                          'combine': True})
    E, _ = eng.run()
    return {'E': float(E), 'S_mid': float(psi.entanglement_entropy()[9]), 'chi': [int(c) for c in psi.chi]}


def case_tfi_tebd_imag(dropin):
    """imaginary-time TEBD of the reference (tebd.py:446 update_bond) towards the ground state"""
    from tenpy.models.tf_ising import TFIChain
    from tenpy.networks.mps import MPS
    from tenpy.algorithms import tebd
    L = 10
    M = TFIChain({'L': L, 'J': 1., 'g': 1.5, 'bc_MPS': 'finite', 'conserve': None})
    psi = MPS.from_product_state(M.lat.mps_sites(), ['up'] * L, bc='finite')
    eng = tebd.TEBDEngine(psi, M, {'order': 2, 'delta_tau_list': [0.1, 0.01], 'N_steps': 5, 'max_error_E': 1.e-6,
                                   'trunc_params': {'chi_max': 20, 'svd_min': 1.e-10}})
    eng.run_GS()
    E = M.bond_energies(psi)
    return {'E': float(sum(E)), 'S_mid': float(psi.entanglement_entropy()[L // 2 - 1]), 'chi': [int(c) for c in psi.chi]}


CASES = {'tfi_dmrg': case_tfi_dmrg, 'xxz_dmrg_mixer': case_xxz_dmrg_mixer, 'tfi_dmrg_fast_engine': case_tfi_dmrg_fast_engine,
         'tfi_tebd_imag': case_tfi_tebd_imag}


def main():
    mode = sys.argv[1]
    names = sys.argv[2:] or list(CASES)
    dropin = setup(mode)
    out = {}
    for name in names:
        out[name] = CASES[name](dropin)
        print(json.dumps({name: out[name]}))
        sys.stdout.flush()
    if mode == 'golden':
        with open(os.path.join(ROOT, 'tests', 'golden', 'dropin.json'), 'w') as f:
            json.dump(out, f, indent=1, sort_keys=True)


if __name__ == '__main__':
    main()
