"""TEBD bond updates against golden vectors of the unmodified reference
(tests/golden/make_golden_tebd.py -> tests/golden/tebd.npz).

The same checks run twice: on the numpy test double (host logic, runs without a GPU) and, marked ``gpu``, on the
CUDA path through the C ABI.  Tolerances: bond energies and entropies 1e-9 absolute after up to 60 sweeps of
non-unitary updates (sums of O(100) bond updates, each 1e-13 accurate), norms 1e-9 relative, chi exact."""
import numpy as np
import pytest

import helpers as h


def _cases():
    from tenpy_b200.models import TFIChain, SpinChain, FermiHubbardChain
    L = 10
    yield 'tfi', TFIChain({'L': L, 'J': 1., 'g': 1.2, 'conserve': None}), ['up'] * L
    yield 'tfip', TFIChain({'L': L, 'J': 1., 'g': 0.8, 'conserve': 'parity'}), ['up'] * L
    yield 'xxz', SpinChain({'L': L, 'Jx': 1., 'Jy': 1., 'Jz': 1.3, 'conserve': 'Sz'}), ['up', 'down'] * (L // 2)
    L = 6
    yield 'hub', FermiHubbardChain({'L': L, 't': 1., 'U': 3., 'mu': 0.}), ['up', 'down'] * (L // 2)


def _compare(tag, g, M, psi, eng, tol):
    Eb = M.bond_energies(psi)
    assert np.max(np.abs(Eb - g[tag + '_Ebond'])) < tol, (tag, np.max(np.abs(Eb - g[tag + '_Ebond'])))
    S = psi.entanglement_entropy()
    assert np.max(np.abs(S - g[tag + '_S'])) < tol * 10, (tag, np.max(np.abs(S - g[tag + '_S'])))
    assert list(psi.chi) == list(g[tag + '_chi']), tag
    assert abs(psi.norm - g[tag + '_norm']) < 1e-9 * abs(g[tag + '_norm']), tag
    assert abs(eng.trunc_err.eps - g[tag + '_eps']) < 1e-12 + 1e-6 * abs(g[tag + '_eps']), tag


def _run_all(names, tol=1e-9):
    from tenpy_b200.networks.mps import MPS
    from tenpy_b200.algorithms.tebd import TEBDEngine
    g = h.load('tebd.npz')
    for name, M, state in _cases():
        if name not in names:
            continue
        L = M.L
        # operators: H_bond (p0, p0*, p1, p1*) and the half-step U_bond (p0, p1, p0*, p1*) equal the reference's
        assert np.max(np.abs(M.H_bond[L // 2].to_ndarray() - g[name + '_Hbond_mid'])) < 1e-14
        assert np.max(np.abs(M.H_bond[1].to_ndarray() - g[name + '_Hbond_first'])) < 1e-14
        psi = MPS.from_product_state(M.lat_sites, state)
        eng = TEBDEngine(psi, M, {'trunc_params': {'chi_max': 32, 'svd_min': 1e-8}})
        eng.calc_U(2, 0.05, type_evo='imag')
        assert np.max(np.abs(eng._U[0][L // 2].to_ndarray() - g[name + '_U_half_mid'])) < 1e-14
        eng.update_imag(30)
        _compare(name + '_imag', g, M, psi, eng, tol)
        assert np.nanmax(psi.isometry_test()) < 1e-11               # the sweeps keep the canonical form
        for order in (1, 2, 4):
            psi = MPS.from_product_state(M.lat_sites, state)
            eng = TEBDEngine(psi, M, {'trunc_params': {'chi_max': 32, 'svd_min': 1e-8}})
            eng.calc_U(order, 0.02, type_evo='imag')
            eng.evolve(6, 0.02)
            _compare('{0}_o{1}'.format(name, order), g, M, psi, eng, tol)


def _run_qr_based(tol=1e-9):
    """QRBasedTEBDEngine (reference tebd.py:619) against the reference: fixed numbers of imaginary-time steps through
    update_bond_imag (sweeps) and update_bond (brick wall); the truncation errors are at rounding level here (1e-15,
    they are sums of differences of O(1) norms), so they are compared absolutely"""
    from tenpy_b200.models import TFIChain, SpinChain
    from tenpy_b200.networks.mps import MPS
    from tenpy_b200.algorithms.tebd import QRBasedTEBDEngine
    g = h.load('tebd_qr.npz')
    opts = {'trunc_params': {'chi_max': 24, 'svd_min': 1e-8}, 'cbe_expand': 0.1, 'cbe_expand_0': 0.5,
            'cbe_min_block_increase': 2, 'compute_err': True}
    L = 10
    for name, M, state in (('tfi', TFIChain({'L': L, 'J': 1., 'g': 1.2, 'conserve': None}), ['up'] * L),
                           ('xxz', SpinChain({'L': L, 'Jx': 1., 'Jy': 1., 'Jz': 1.3, 'conserve': 'Sz'}),
                            ['up', 'down'] * (L // 2))):
        for tag, run in (('_imag', lambda e: (e.calc_U(2, 0.05, type_evo='imag'), e.update_imag(20))),
                         ('_o2', lambda e: (e.calc_U(2, 0.02, type_evo='imag'), e.evolve(6, 0.02)))):
            psi = MPS.from_product_state(M.lat_sites, state)
            eng = QRBasedTEBDEngine(psi, M, dict(opts))
            run(eng)
            t = name + tag
            Eb = M.bond_energies(psi)
            assert np.max(np.abs(Eb - g[t + '_Ebond'])) < tol, (t, np.max(np.abs(Eb - g[t + '_Ebond'])))
            assert np.max(np.abs(psi.entanglement_entropy() - g[t + '_S'])) < tol * 10, t
            assert list(psi.chi) == list(g[t + '_chi']), (t, psi.chi, g[t + '_chi'])
            assert abs(psi.norm - g[t + '_norm']) < 1e-9 * abs(g[t + '_norm']), t
            assert abs(eng.trunc_err.eps - g[t + '_eps']) < 1e-13, t


def test_tebd_qr_based_host_logic(fake_device):
    _run_qr_based()


@pytest.mark.gpu
def test_tebd_qr_based_gpu(gpu_lib):
    _run_qr_based()


def test_trotter_schedules():
    from tenpy_b200.algorithms.tebd import TEBDEngine as T
    assert T.suzuki_trotter_decomposition(2, 3) == [(0, 1), (1, 0), (1, 1), (1, 0), (1, 1), (1, 0), (0, 1)]
    for order in (1, 2, 4):
        dts = T.suzuki_trotter_time_steps(order)
        for N in (1, 3):
            tot = [0., 0.]
            for j, k in T.suzuki_trotter_decomposition(order, N):
                tot[k] += dts[j]
            assert abs(tot[0] - N) < 1e-14 and abs(tot[1] - N) < 1e-14   # every layer family adds up to N dt
    with pytest.raises(NotImplementedError):
        T(type('P', (), {'L': 2})(), None, {}).calc_U(2, 0.1, 'real')


def test_tebd_host_logic(fake_device):
    _run_all(['tfi', 'tfip', 'xxz', 'hub'])


@pytest.mark.gpu
def test_tebd_gpu_parity(gpu_lib):
    _run_all(['tfi', 'tfip', 'xxz', 'hub'])


def _run_GS_check():
    """imaginary-time TEBD ground state against DMRG, as the reference's tests/test_tebd.py:73-90: `run_GS`, then
    `psi.canonical_form()` (imaginary time evolution leaves the canonical form), then the sum of bond energies."""
    from tenpy_b200.models import TFIChain
    from tenpy_b200.networks.mps import MPS
    from tenpy_b200.algorithms.tebd import TEBDEngine
    from tenpy_b200.algorithms import dmrg
    L = 10
    M = TFIChain({'L': L, 'J': 1., 'g': 1.2, 'conserve': None})
    psi = MPS.from_product_state(M.lat_sites, ['up'] * L)
    eng = TEBDEngine(psi, M, {'trunc_params': {'chi_max': 32, 'svd_min': 1e-8}, 'delta_tau_list': [0.1, 0.01, 0.001],
                              'max_error_E': 1e-10, 'N_steps': 10})
    eng.run_GS()
    psi.canonical_form()
    E_tebd = np.sum(M.bond_energies(psi))
    psi2 = MPS.from_product_state(M.lat_sites, ['up'] * L)
    res = dmrg.run(psi2, M, {'mixer': None, 'max_E_err': 1e-11, 'trunc_params': {'chi_max': 32, 'svd_min': 1e-8}})
    assert abs((E_tebd - res['E']) / res['E']) < 1e-7, (E_tebd, res['E'])     # Trotter error O(dtau^2) at dtau=1e-3


def test_tebd_run_GS_host_logic(fake_device):
    _run_GS_check()


@pytest.mark.gpu
def test_tebd_run_GS_matches_dmrg(gpu_lib):
    _run_GS_check()
