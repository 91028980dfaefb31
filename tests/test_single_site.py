"""Single-site DMRG: OneSiteH.matvec against the reference's own tensors, the engine against
the reference's energies / entropies (tests/golden/dmrg_1site.npz from tests/golden/make_golden_1site.py)."""
import numpy as np
import pytest

import helpers as h


def _check_matvec():
    from tenpy_b200.linalg import np_conserved as npc
    g = h.load('dmrg_1site.npz')
    # right move: LHeff . theta . RP ; left move: LP . theta . RHeff  -- the reference's tensors of a converged XXZ chain
    LHeff, RP, th = (h.to_product(h.oarray_from(g, 'xxz_r_' + k)) for k in ('LHeff', 'RP', 'theta'))
    t = npc.tensordot(LHeff, th, axes=['(vR.p0*)', '(vL.p0)'])
    t = npc.tensordot(t, RP, axes=[['wR', 'vR'], ['wL', 'vL']])
    t.ireplace_labels(['(vR*.p0)', 'vL*'], ['(vL.p0)', 'vR'])
    h.assert_close(h.to_oracle(t.itranspose(['(vL.p0)', 'vR'])), h.oarray_from(g, 'xxz_r_Htheta'), 1e-13)
    RHeff, LP, th = (h.to_product(h.oarray_from(g, 'xxz_l_' + k)) for k in ('RHeff', 'LP', 'theta'))
    t = npc.tensordot(th, RHeff, axes=['(p0.vR)', '(p0*.vL)'])
    t = npc.tensordot(LP, t, axes=[['vR', 'wR'], ['vL', 'wL']])
    t.ireplace_labels(['vR*', '(p0.vL*)'], ['vL', '(p0.vR)'])
    h.assert_close(h.to_oracle(t.itranspose(['vL', '(p0.vR)'])), h.oarray_from(g, 'xxz_l_Htheta'), 1e-13)


def _check_onesite_H_consistency():
    """OneSiteH: combine=True (both directions) and combine=False give the same linear map, equal to to_matrix()"""
    from tenpy_b200.models import SpinChain
    from tenpy_b200.networks.mps import MPS
    from tenpy_b200.algorithms import dmrg
    from tenpy_b200.algorithms.mps_common import OneSiteH
    from tenpy_b200.linalg import np_conserved as npc
    L = 8
    M = SpinChain({'L': L, 'Jx': 1., 'Jy': 1., 'Jz': 0.8, 'conserve': 'Sz'})
    psi = MPS.from_product_state(M.lat_sites, ['up', 'down'] * (L // 2))
    eng = dmrg.TwoSiteDMRGEngine(psi, M, {'mixer': True, 'trunc_params': {'chi_max': 12, 'svd_min': 1e-12}})
    eng.sweep()
    eng.sweep()
    for i0 in (2, 4):
        theta = psi.get_theta(i0, 1)
        ref = OneSiteH(eng.env, i0, combine=False).matvec(theta)
        for move_right in (True, False):
            H = OneSiteH(eng.env, i0, combine=True, move_right=move_right)
            got = H.matvec(H.combine_theta(theta)).split_legs()
            got.itranspose(ref.get_leg_labels())
            assert npc.norm(got - ref) < 1e-13 * npc.norm(ref)
            mat = H.to_matrix().to_ndarray()
            assert np.max(np.abs(mat - mat.T)) < 1e-12
            vec = H.combine_theta(theta).combine_legs(H.acts_on, qconj=+1)
            assert np.max(np.abs(mat @ vec.to_ndarray() -
                                 H.matvec(H.combine_theta(theta)).combine_legs(H.acts_on, qconj=+1).to_ndarray())) < 1e-12


def _check_engine():
    from tenpy_b200.models import TFIChain, SpinChain
    from tenpy_b200.networks.mps import MPS
    from tenpy_b200.algorithms import dmrg
    g = h.load('dmrg_1site.npz')
    # (1) TFI from a product state, DensityMatrixMixer working on the two-site theta
    for combine in (True, False):
        L = 12
        M = TFIChain({'L': L, 'J': 1., 'g': 1.1, 'conserve': None})
        psi = MPS.from_product_state(M.lat_sites, ['up'] * L)
        res = dmrg.run(psi, M, {'active_sites': 1, 'mixer': 'DensityMatrixMixer',
                                'mixer_params': {'amplitude': 1e-3, 'decay': 2., 'disable_after': 8},
                                'max_E_err': 1e-11, 'max_S_err': 1e-8, 'trunc_params': {'chi_max': 24, 'svd_min': 1e-10},
                                'combine': combine, 'max_sweeps': 24})
        key = 'tfi_c%d' % int(combine)
        assert abs(res['E'] - g[key + '_E']) < 1e-10 * abs(g[key + '_E'])
        assert np.max(np.abs(psi.entanglement_entropy() - g[key + '_S'])) < 1e-7
        assert np.max(psi.isometry_test()) < 1e-11
    # (2) XXZ with Sz
    L = 10
    M = SpinChain({'L': L, 'Jx': 1., 'Jy': 1., 'Jz': 0.8, 'conserve': 'Sz'})
    psi = MPS.from_product_state(M.lat_sites, ['up', 'down'] * (L // 2))
    res = dmrg.run(psi, M, {'active_sites': 1, 'mixer': 'DensityMatrixMixer',
                            'mixer_params': {'amplitude': 1e-3, 'decay': 2., 'disable_after': 8},
                            'max_E_err': 1e-11, 'max_S_err': 1e-8, 'trunc_params': {'chi_max': 32, 'svd_min': 1e-10},
                            'combine': True, 'max_sweeps': 24})
    assert abs(res['E'] - g['xxz_E']) < 1e-10 * abs(g['xxz_E'])
    assert np.max(np.abs(psi.entanglement_entropy() - g['xxz_S'])) < 1e-7
    # (3) mixer-free single-site refinement of a truncated two-site state
    L = 14
    M = TFIChain({'L': L, 'J': 1., 'g': 0.9, 'conserve': None})
    psi = MPS.from_product_state(M.lat_sites, ['up'] * L)
    e2 = dmrg.TwoSiteDMRGEngine(psi, M, {'mixer': None, 'trunc_params': {'chi_max': 8, 'svd_min': 1e-12},
                                         'max_sweeps': 2, 'min_sweeps': 2, 'combine': True})
    e2.run()
    assert abs(e2.sweep_stats['E'][-1] - g['ref_E2']) < 1e-10 * abs(g['ref_E2'])
    res = dmrg.run(psi, M, {'active_sites': 1, 'mixer': None, 'trunc_params': {'chi_max': 8, 'svd_min': 1e-12},
                            'max_E_err': 1e-12, 'max_sweeps': 12, 'combine': True})
    assert abs(res['E'] - g['ref_E1']) < 1e-10 * abs(g['ref_E1'])
    assert res['E'] <= g['ref_E2'] + 1e-12
    assert np.max(np.abs(psi.entanglement_entropy() - g['ref_S1'])) < 1e-7


def _check_subspace_expansion():
    """the reference's default one-site mixer: single-site engine (combine on / off) and the two-site engine through
    Mixer.mix_and_decompose_2site, energies / entropies of the reference"""
    from tenpy_b200.models import TFIChain, SpinChain
    from tenpy_b200.networks.mps import MPS
    from tenpy_b200.algorithms import dmrg
    from tenpy_b200.algorithms.mps_common import SubspaceExpansion
    g = h.load('dmrg_1site.npz')
    mp = {'amplitude': 1e-3, 'decay': 2., 'disable_after': 8}
    L = 12
    M = TFIChain({'L': L, 'J': 1., 'g': 1.1, 'conserve': None})
    psi = MPS.from_product_state(M.lat_sites, ['up'] * L)
    eng = dmrg.SingleSiteDMRGEngine(psi, M, {'mixer': True, 'mixer_params': dict(mp), 'max_E_err': 1e-11,
                                             'max_S_err': 1e-8, 'trunc_params': {'chi_max': 24, 'svd_min': 1e-10},
                                             'combine': True, 'max_sweeps': 24})
    E, _ = eng.run()
    assert eng.DefaultMixer is SubspaceExpansion
    assert abs(E - g['se_tfi_E']) < 1e-10 * abs(g['se_tfi_E'])
    assert np.max(np.abs(psi.entanglement_entropy() - g['se_tfi_S'])) < 1e-7
    assert np.max(psi.isometry_test()) < 1e-11
    L = 10
    M = SpinChain({'L': L, 'Jx': 1., 'Jy': 1., 'Jz': 0.8, 'conserve': 'Sz'})
    for key, Engine, combine in (('se_xxz1', dmrg.SingleSiteDMRGEngine, True), ('se_xxz1n', dmrg.SingleSiteDMRGEngine, False),
                                 ('se_xxz2', dmrg.TwoSiteDMRGEngine, True)):
        psi = MPS.from_product_state(M.lat_sites, ['up', 'down'] * (L // 2))
        eng = Engine(psi, M, {'mixer': 'SubspaceExpansion', 'mixer_params': dict(mp), 'max_E_err': 1e-11,
                              'max_S_err': 1e-8, 'trunc_params': {'chi_max': 32, 'svd_min': 1e-10},
                              'combine': combine, 'max_sweeps': 24})
        E, _ = eng.run()
        assert abs(E - g[key + '_E']) < 1e-10 * abs(g[key + '_E']), key
        assert np.max(np.abs(psi.entanglement_entropy() - g[key + '_S'])) < 1e-7, key
        assert list(psi.chi) == list(g[key + '_chi']), key


def _check_mix_both_sides():
    """Mixer.mix_and_decompose_2site with both bonds mixed (reference mps_common.py:1770-1785; the infinite-DMRG case):
    both factors are isometries, U . S . VH reproduces theta up to the truncation"""
    from tenpy_b200.models import SpinChain
    from tenpy_b200.networks.mps import MPS
    from tenpy_b200.algorithms import dmrg
    from tenpy_b200.algorithms.mps_common import SubspaceExpansion
    from tenpy_b200.linalg import np_conserved as npc
    L = 8
    M = SpinChain({'L': L, 'Jx': 1., 'Jy': 1., 'Jz': 0.8, 'conserve': 'Sz'})
    psi = MPS.from_product_state(M.lat_sites, ['up', 'down'] * (L // 2))
    eng = dmrg.TwoSiteDMRGEngine(psi, M, {'mixer': True, 'trunc_params': {'chi_max': 40, 'svd_min': 1e-14}})
    eng.sweep()
    eng.sweep()
    eng.i0, eng.move_right, eng.update_LP_RP = 3, True, (True, True)
    theta = eng.prepare_update_local()
    mixer = SubspaceExpansion({'amplitude': 1e-6})
    qL = psi.get_B(3, form=None).qtotal
    U, S, VH, err, S_a = mixer.mix_and_decompose_2site(eng, theta, 3, True, True, [qL, theta.qtotal - qL])
    iso = npc.tensordot(U.conj(), U, axes=['(vL*.p0*)', '(vL.p0)']).to_ndarray()
    assert np.max(np.abs(iso - np.eye(len(iso)))) < 1e-11
    iso = npc.tensordot(VH, VH.conj(), axes=['(p1.vR)', '(p1*.vR*)']).to_ndarray()
    assert np.max(np.abs(iso - np.eye(len(iso)))) < 1e-11
    rec = npc.tensordot(npc.tensordot(U, S, axes=['vR', 'vL']), VH, axes=['vR', 'vL'])
    ov = abs(npc.inner(rec, theta, axes='range', do_conj=True)) / (npc.norm(rec) * npc.norm(theta))
    assert abs(ov - 1.) < 1e-9 and abs(npc.norm(S) - 1.) < 1e-12


def test_subspace_expansion_host_logic(fake_device):
    _check_mix_both_sides()
    _check_subspace_expansion()


@pytest.mark.gpu
def test_subspace_expansion_gpu(gpu_lib):
    _check_subspace_expansion()


def test_onesite_matvec_host_logic(fake_device):
    _check_matvec()
    _check_onesite_H_consistency()


def test_single_site_engine_host_logic(fake_device):
    _check_engine()


@pytest.mark.gpu
def test_onesite_matvec_gpu(gpu_lib):
    _check_matvec()
    _check_onesite_H_consistency()


@pytest.mark.gpu
def test_single_site_engine_gpu(gpu_lib):
    _check_engine()
