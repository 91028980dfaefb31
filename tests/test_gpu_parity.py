"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle and the golden vectors of the
reference, on the same inputs.

Tolerances (written next to each assert): block tables / charges / slices are integer work and must be
IDENTICAL; floating point blocks agree with the oracle to 1e-13 * scale (FP64 with a different summation
order); ground-state energy and entanglement entropy to 1e-10 relative and singular values to 1e-8
(BASELINE.json north_star).  At the full benchmark sizes the oracle is too slow, so size-independent
properties are checked instead (linearity and hermiticity of the matvec, U S VH reconstruction,
isometry, combine/split round trips).
"""
import numpy as np
import pytest

import helpers as h
from oracle import npc_blocks as ob
from oracle import dmrg_dense as od

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def g_td():
    return h.load('tensordot.npz')


@pytest.fixture(scope='module')
def g_rs():
    return h.load('reshape_svd.npz')


@pytest.fixture(scope='module')
def g_dm():
    return h.load('dmrg.npz')


def test_tensordot_inner_norm_axpy(gpu_lib, g_td):
    from tenpy_b200.linalg import np_conserved as npc
    for ci in range(int(g_td['ncases'])):
        oa, obb, oc = (h.oarray_from(g_td, 'c%d_%s' % (ci, k)) for k in 'abc')
        a, b = h.to_product(oa), h.to_product(obb)
        n = int(g_td['c%d_naxes' % ci])
        c = npc.tensordot(a, b, axes=n)
        c.test_sanity()
        h.assert_close(h.to_oracle(c), oc, 1e-13)                       # vs reference (golden)
        h.assert_close(h.to_oracle(c), ob.tensordot(oa, obb, n), 1e-13)  # vs oracle
        assert np.array_equal(h.to_oracle(a).to_dense(), oa.to_dense()), 'inputs must not be modified'
        assert abs(npc.inner(a, a, 'range', do_conj=True) - g_td['c%d_inner_aa' % ci]) < 1e-12
        assert abs(npc.norm(a) - g_td['c%d_norm_a' % ci]) < 1e-13
        a2 = h.to_product(h.oarray_from(g_td, 'c%d_a2' % ci))
        assert abs(npc.inner(a, a2, 'range', do_conj=True) - g_td['c%d_inner_aa2' % ci]) < 1e-12
        h.assert_close(h.to_oracle(a + a2 * 0.37), h.oarray_from(g_td, 'c%d_sum' % ci), 1e-14)
        a3 = a.copy()
        a3.iscale_prefactor(-2.5)
        assert np.max(np.abs(a3.to_ndarray() + 2.5 * oa.to_dense())) < 1e-15 * 10


def test_block_moves_exact(gpu_lib, g_rs):
    """combine_legs / split_legs / transpose / iproject are pure data movement: bit exact"""
    a = h.to_product(h.oarray_from(g_rs, 'a'))
    comb = a.combine_legs([['vL', 'p0'], ['p1', 'vR']], qconj=[+1, -1])
    h.assert_close(h.to_oracle(comb), h.oarray_from(g_rs, 'comb'), 0.)
    comb2 = a.combine_legs([['vR', 'p1'], ['p0', 'vL']], new_axes=[0, 2], qconj=[-1, +1])
    h.assert_close(h.to_oracle(comb2), h.oarray_from(g_rs, 'comb2'), 0.)
    h.assert_close(h.to_oracle(comb.split_legs()), h.oarray_from(g_rs, 'split'), 0.)
    h.assert_close(h.to_oracle(a.transpose(['p1', 'vL', 'w', 'vR', 'p0'])), h.oarray_from(g_rs, 'transp'), 0.)
    back = comb.split_legs().transpose(['vL', 'p0', 'w', 'p1', 'vR'])
    assert np.array_equal(back.to_ndarray(), a.to_ndarray())
    m = h.to_product(h.oarray_from(g_rs, 'm'))
    mp = m.copy()
    mp.iproject(g_rs['proj_mask'], 1)
    h.assert_close(h.to_oracle(mp), h.oarray_from(g_rs, 'm_proj'), 0.)
    h.assert_close(h.to_oracle(m.scale_axis(g_rs['scale_s'], 0)), h.oarray_from(g_rs, 'm_scaled'), 1e-15)


def test_svd_eigh_vs_reference(gpu_lib, g_rs):
    from tenpy_b200.linalg import np_conserved as npc
    from tenpy_b200.linalg.truncation import svd_theta
    m = h.to_product(h.oarray_from(g_rs, 'm'))
    U, S, VH = npc.svd(m, inner_labels=['vR', 'vL'])
    h.assert_same_structure(h.to_oracle(U), h.oarray_from(g_rs, 'm_U'))       # identical block tables / new leg
    h.assert_same_structure(h.to_oracle(VH), h.oarray_from(g_rs, 'm_VH'))
    assert np.max(np.abs(S - g_rs['m_S'])) < 1e-8 * np.max(S)                  # block-ordered like the reference
    assert np.max(np.abs(S - g_rs['m_S'])) < 1e-12
    rec = npc.tensordot(U.scale_axis(S, 1), VH, axes=1)
    assert np.max(np.abs(rec.to_ndarray() - m.to_ndarray())) < 1e-13
    UdU = npc.tensordot(U.conj(), U, axes=[0, 0]).to_ndarray()
    assert np.max(np.abs(UdU - np.eye(len(S)))) < 1e-13
    VVd = npc.tensordot(VH, VH.conj(), axes=[1, 1]).to_ndarray()
    assert np.max(np.abs(VVd - np.eye(len(S)))) < 1e-13
    Ut, St, VHt, err, renorm = svd_theta(m, {'chi_max': 17, 'svd_min': 1e-8}, inner_labels=['vR', 'vL'])
    h.assert_same_structure(h.to_oracle(Ut), h.oarray_from(g_rs, 'm_Ut'))
    assert np.max(np.abs(St - g_rs['m_St'])) < 1e-12 and abs(err.eps - g_rs['m_err']) < 1e-14
    rho = h.to_product(h.oarray_from(g_rs, 'rho'))
    w, V = npc.eigh(rho)
    h.assert_same_structure(h.to_oracle(V), h.oarray_from(g_rs, 'rho_V'))
    assert np.max(np.abs(w - g_rs['rho_w'])) < 1e-12 * max(1., np.max(np.abs(g_rs['rho_w'])))
    Vd = V.to_ndarray()
    assert np.max(np.abs(Vd @ np.diag(w) @ Vd.T - rho.to_ndarray())) < 1e-12
    assert np.max(np.abs(Vd.T @ Vd - np.eye(len(w)))) < 1e-13


def test_two_site_matvec_lanczos_env(gpu_lib, g_dm):
    """the hot path on the converged XXZ (Sz-conserving) state of the reference: matvec, Lanczos, env update"""
    from tenpy_b200.linalg import np_conserved as npc
    from tenpy_b200.linalg.krylov_based import LanczosGroundState
    oL, oR, oT = (h.oarray_from(g_dm, 'xxz_' + k) for k in ('LHeff', 'RHeff', 'theta'))
    LHeff, RHeff, theta = h.to_product(oL), h.to_product(oR), h.to_product(oT)

    class H:
        def matvec(self, th):
            labels = th.get_leg_labels()
            t = npc.tensordot(LHeff, th, axes=['(vR.p0*)', '(vL.p0)'])
            t = npc.tensordot(t, RHeff, axes=[['wR', '(p1.vR)'], ['wL', '(p1*.vL)']])
            t.ireplace_labels(['(vR*.p0)', '(p1.vL*)'], ['(vL.p0)', '(p1.vR)'])
            return t.itranspose(labels)
    Hth = H().matvec(theta)
    h.assert_close(h.to_oracle(Hth), h.oarray_from(g_dm, 'xxz_Htheta'), 1e-13)
    h.assert_close(h.to_oracle(Hth), ob.two_site_matvec(oL, oR, oT), 1e-13)
    E0, th0, N = LanczosGroundState(H(), theta, {}).run()
    assert abs(E0 - g_dm['xxz_lanczos_E0']) < 1e-10 * abs(g_dm['xxz_lanczos_E0'])
    assert abs(E0 - g_dm['xxz_E']) < 1e-10 * abs(g_dm['xxz_E'])
    U, S, VH = npc.svd(th0, inner_labels=['vR', 'vL'])
    k = min(len(S), len(g_dm['xxz_theta_S']))
    assert np.max(np.abs(np.sort(S)[::-1][:k] - np.sort(g_dm['xxz_theta_S'])[::-1][:k])) < 1e-8
    Ug = h.to_product(h.oarray_from(g_dm, 'xxz_U'))
    LP = npc.tensordot(LHeff, Ug, axes=['(vR.p0*)', '(vL.p0)'])
    LP = npc.tensordot(Ug.conj(), LP, axes=['(vL*.p0*)', '(vR*.p0)'])
    h.assert_close(h.to_oracle(LP), h.oarray_from(g_dm, 'xxz_LPnew'), 1e-12)


def _run(model, p_state, opts):
    from tenpy_b200.networks.mps import MPS
    from tenpy_b200.algorithms import dmrg
    psi = MPS.from_product_state(model.lat_sites, p_state)
    return dmrg.run(psi, model, opts), psi


def _check(res, psi, g, key, L):
    E = g[key + '_E']
    assert abs(res['E'] - E) < 1e-10 * abs(E), (res['E'], E)                   # energy: 1e-10 relative
    S = psi.entanglement_entropy()
    assert np.max(np.abs(S - g[key + '_S'])) < 1e-10 * 100, np.max(np.abs(S - g[key + '_S']))  # entropy
    if key + '_sv_mid' in g:
        sv = np.sort(np.asarray(psi.get_SL(L // 2)))[::-1]
        ref = np.sort(g[key + '_sv_mid'])[::-1]
        k = min(len(sv), len(ref))
        assert np.max(np.abs(sv[:k] - ref[:k])) < 1e-8                          # singular values: 1e-8
    assert np.nanmax(psi.isometry_test()) < 1e-11


def test_dmrg_config1_tfi(gpu_lib, g_dm):
    """BASELINE.json configs[0]: TFIChain L=20 chi=50; reference E = -25.1077971116238"""
    from tenpy_b200.models import TFIChain
    M = TFIChain({'L': 20, 'J': 1., 'g': 1., 'conserve': None})
    res, psi = _run(M, ['up'] * 20, {'mixer': None, 'max_E_err': 1e-10, 'combine': True,
                                    'trunc_params': {'chi_max': 50, 'svd_min': 1e-10}})
    _check(res, psi, g_dm, 'tfi', 20)
    assert abs(res['E'] - (-25.1077971116238)) < 2.6e-9
    ref = od.run_dmrg(od.tfi_mpo(1., 1.), 20, 2, [0] * 20, dict(chi_max=50, svd_min=1e-10), {}, max_E_err=1e-10)
    assert abs(res['E'] - ref['E']) < 1e-10 * abs(ref['E'])                     # vs the dense CPU oracle


def test_dmrg_xxz_sz_mixer(gpu_lib, g_dm):
    """block-sparse path with the density-matrix mixer (small version of configs[2])"""
    from tenpy_b200.models import SpinChain
    L = 16
    M = SpinChain({'L': L, 'Jx': 1., 'Jy': 1., 'Jz': 1., 'conserve': 'Sz'})
    opts = {'mixer': True, 'mixer_params': {'amplitude': 1e-5, 'decay': 2., 'disable_after': 6}, 'max_E_err': 1e-11,
            'max_S_err': 1e-8, 'trunc_params': {'chi_max': 60, 'svd_min': 1e-10}, 'combine': True, 'max_sweeps': 20}
    res, psi = _run(M, ['up', 'down'] * (L // 2), opts)
    _check(res, psi, g_dm, 'xxz', L)


def test_dmrg_hubbard_n_sz(gpu_lib, g_dm):
    """two conserved charges (N, Sz), many small blocks (small version of configs[3])"""
    from tenpy_b200.models import FermiHubbardChain
    L = 6
    M = FermiHubbardChain({'L': L, 't': 1., 'U': 4., 'mu': 0.})
    opts = {'mixer': True, 'mixer_params': {'amplitude': 1e-5, 'decay': 2., 'disable_after': 6}, 'max_E_err': 1e-11,
            'max_S_err': 1e-8, 'trunc_params': {'chi_max': 64, 'svd_min': 1e-10}, 'combine': True, 'max_sweeps': 20}
    res, psi = _run(M, ['up', 'down'] * (L // 2), opts)
    _check(res, psi, g_dm, 'hub', L)


def test_dmrg_tfi_parity(gpu_lib, g_dm):
    from tenpy_b200.models import TFIChain
    M = TFIChain({'L': 12, 'J': 1., 'g': 0.8, 'conserve': 'parity'})
    res, psi = _run(M, ['up'] * 12, {'mixer': True, 'mixer_params': {'disable_after': 5}, 'max_E_err': 1e-11,
                                    'trunc_params': {'chi_max': 40, 'svd_min': 1e-10}, 'combine': True,
                                    'max_sweeps': 16})
    _check(res, psi, g_dm, 'tfip', 12)


_MID_OPTS = {'mixer': True, 'mixer_params': {'amplitude': 1e-5, 'decay': 2., 'disable_after': 8}, 'max_E_err': 1e-11,
             'max_S_err': 1e-10, 'combine': True, 'max_sweeps': 40,
             'lanczos_params': {'P_tol': 1e-22, 'N_max': 40}}


def test_dmrg_mid_xxz_sz(gpu_lib):
    """scaled-down BASELINE.json configs[2]: XXZ L=32, Sz conserved, chi=96 (truncating); golden from the
    reference (tests/golden/make_golden_mid.py): E = -16.50060013071288"""
    from tenpy_b200.models import SpinChain
    g = h.load('dmrg_mid.npz')
    L = 32
    M = SpinChain({'L': L, 'Jx': 1., 'Jy': 1., 'Jz': 1.5, 'conserve': 'Sz'})
    res, psi = _run(M, ['up', 'down'] * (L // 2), dict(_MID_OPTS, trunc_params={'chi_max': 96, 'svd_min': 1e-10}))
    assert list(psi.chi) == list(g['xxz32_chi'])
    _check(res, psi, g, 'xxz32', L)


def test_dmrg_mid_hubbard(gpu_lib):
    """scaled-down BASELINE.json configs[3]: Fermi-Hubbard L=12, (N, Sz) conserved, chi=160; reference
    E = -6.526243382515468"""
    from tenpy_b200.models import FermiHubbardChain
    g = h.load('dmrg_mid.npz')
    L = 12
    M = FermiHubbardChain({'L': L, 't': 1., 'U': 4., 'mu': 0.})
    res, psi = _run(M, ['up', 'down'] * (L // 2), dict(_MID_OPTS, trunc_params={'chi_max': 1000, 'svd_min': 1e-5}))
    _check(res, psi, g, 'hub12', L)


def test_full_size_properties(gpu_lib):
    """BASELINE.json configs[1] shapes (chi=1024, d=2, D=3): size-independent properties of the hot path"""
    import torch
    from tenpy_b200.linalg import np_conserved as npc
    chi, d, D = 1024, 2, 3
    n = chi * d
    ci = npc.ChargeInfo()
    lL, lR, lW = (npc.LegCharge.from_trivial(n, ci, +1), npc.LegCharge.from_trivial(n, ci, -1),
                  npc.LegCharge.from_trivial(D, ci, -1))
    gen = torch.Generator(device='cuda')
    gen.manual_seed(7)

    def rnd(legs, labels):
        t = torch.randn(int(np.prod([l.ind_len for l in legs])), dtype=torch.float64, device='cuda', generator=gen)
        return npc.Array.from_device_buffer(legs, np.zeros((1, len(legs)), np.int64), t, labels=labels)
    LHeff = rnd([lL, lW, lL.conj()], ['(vR*.p0)', 'wR', '(vR.p0*)'])
    RHeff = rnd([lW.conj(), lR.conj(), lR], ['wL', '(p1*.vL)', '(p1.vL*)'])
    x = rnd([lL, lR], ['(vL.p0)', '(p1.vR)'])
    y = rnd([lL, lR], ['(vL.p0)', '(p1.vR)'])

    def mv(th):
        t = npc.tensordot(LHeff, th, axes=['(vR.p0*)', '(vL.p0)'])
        t = npc.tensordot(t, RHeff, axes=[['wR', '(p1.vR)'], ['wL', '(p1*.vL)']])
        return t.ireplace_labels(['(vR*.p0)', '(p1.vL*)'], ['(vL.p0)', '(p1.vR)'])
    # linearity: H(2x - 3y) = 2Hx - 3Hy
    lhs = mv(x * 2. + y * (-3.))
    rhs = mv(x) * 2. + mv(y) * (-3.)
    scale = npc.norm(rhs)
    assert npc.norm(lhs - rhs) < 1e-13 * scale
    # a spot check of one output row against numpy on the host (first 4 rows of LHeff are enough)
    Lh = LHeff.to_ndarray()[:4]
    ref = np.tensordot(np.tensordot(Lh, x.to_ndarray(), axes=[2, 0]), RHeff.to_ndarray(), axes=[[1, 2], [0, 1]])
    got = mv(x).to_ndarray()[:4]
    assert np.max(np.abs(got - ref)) < 1e-12 * np.max(np.abs(ref))
    # SVD of a 2048 x 2048 theta with a decaying spectrum (DMRG-like: row-graded, nearly orthogonal rows)
    q, _ = torch.linalg.qr(torch.randn(n, n, dtype=torch.float64, device='cuda', generator=gen))
    s = torch.exp(-torch.arange(n, dtype=torch.float64, device='cuda') / 60.)
    th = npc.Array.from_device_buffer([lL, lR], np.zeros((1, 2), np.int64), (s[:, None] * q).reshape(-1).contiguous(),
                                      labels=['(vL.p0)', '(p1.vR)'])
    th = th + mv(th) * (0.05 / np.sqrt(n) / D)
    U, S, VH = npc.svd(th)
    assert np.all(np.diff(S) <= 1e-15)
    rec = npc.tensordot(U.scale_axis(S, 1), VH, axes=1)
    assert npc.norm(rec - th) < 1e-12 * npc.norm(th)
    keep = S > 1e-9 * S[0]
    Uk = U.copy()
    Uk.iproject(keep, 1)
    G = npc.tensordot(Uk.conj(), Uk, axes=[0, 0]).to_ndarray()
    assert np.max(np.abs(G - np.eye(G.shape[0]))) < 1e-11
    Sref = np.linalg.svd(th.to_ndarray(), compute_uv=False)
    assert np.max(np.abs(S - Sref)) < 1e-8 * Sref[0]


@pytest.mark.gpu
def test_split_and_identity_matvec_routes(gpu_lib):
    """The contraction routes the benchmark takes for large blocks, forced here on small ones: 'split' order (LP, W0 W1, RP on
    the split theta), the identity-environment shortcut (LP[IdL] = RP[IdR] = 1 skipped, views / direct write in the dense
    case) -- per bond against the reference's combined sequence, and whole DMRG runs against the reference's energies."""
    from tenpy_b200.models import TFIChain, SpinChain, FermiHubbardChain
    from tenpy_b200.networks.mps import MPS
    from tenpy_b200.algorithms import dmrg
    from tenpy_b200.algorithms.mps_common import TwoSiteH
    from tenpy_b200.linalg import np_conserved as npc
    cases = [(TFIChain({'L': 8, 'J': 1., 'g': 1.1, 'conserve': None}), ['up'] * 8, None),
             (SpinChain({'L': 8, 'Jx': 1., 'Jy': 1., 'Jz': 0.7, 'conserve': 'Sz'}), ['up', 'down'] * 4, True),
             (FermiHubbardChain({'L': 6, 't': 1., 'U': 4., 'mu': 0.}), ['up', 'down'] * 3, True)]
    for M, state, mixer in cases:
        psi = MPS.from_product_state(M.lat_sites, state)
        eng = dmrg.TwoSiteDMRGEngine(psi, M, {'mixer': mixer, 'combine': True, 'matvec_order': 'combined',
                                              'trunc_params': {'chi_max': 24, 'svd_min': 1e-12}})
        eng.sweep()
        eng.sweep()
        eng.mixer_cleanup()
        psi.canonical_form()
        eng.env.clear()
        used = 0
        for i0 in range(psi.L - 1):
            Hc = TwoSiteH(eng.env, i0, combine=True, matvec_order='combined')
            Hs = TwoSiteH(eng.env, i0, combine=True, matvec_order='split')
            Hs.identity_env = False
            Hi = TwoSiteH(eng.env, i0, combine=True, matvec_order='split')
            Hi.identity_env = True
            theta = Hc.combine_theta(psi.get_theta(i0, 2))
            before = theta.to_ndarray().copy()
            a, b, c = Hc.matvec(theta), Hs.matvec(theta), Hi.matvec(theta)
            c2 = Hi.matvec(theta)                      # second call: cached structures / direct-write path
            scale = max(npc.norm(a), 1e-300)
            assert npc.norm(a - b) <= 1e-13 * scale
            assert npc.norm(a - c) <= 1e-11 * scale and npc.norm(c - c2) <= 1e-14 * scale
            assert np.array_equal(theta.to_ndarray(), before)       # the views never write into the input
            used += int(bool(Hi._id_env))
        assert used >= 1        # (how often it applies on the device is reported by bench.py: identity_env_stats)
    g = h.load('dmrg.npz')
    M = TFIChain({'L': 20, 'J': 1., 'g': 1., 'conserve': None})
    psi = MPS.from_product_state(M.lat_sites, ['up'] * 20)
    res = dmrg.run(psi, M, {'mixer': None, 'max_E_err': 1e-10, 'combine': True, 'matvec_order': 'split',
                            'identity_env': True, 'trunc_params': {'chi_max': 50, 'svd_min': 1e-10}})
    assert abs(res['E'] - g['tfi_E']) < 1e-10 * abs(g['tfi_E'])
    assert np.max(np.abs(psi.entanglement_entropy() - g['tfi_S'])) < 1e-8
    L = 16
    M = SpinChain({'L': L, 'Jx': 1., 'Jy': 1., 'Jz': 1., 'conserve': 'Sz'})
    psi = MPS.from_product_state(M.lat_sites, ['up', 'down'] * (L // 2))
    res = dmrg.run(psi, M, {'mixer': True, 'mixer_params': {'amplitude': 1e-5, 'decay': 2., 'disable_after': 6},
                            'max_E_err': 1e-11, 'max_S_err': 1e-8, 'trunc_params': {'chi_max': 60, 'svd_min': 1e-10},
                            'combine': True, 'max_sweeps': 20, 'matvec_order': 'split', 'identity_env': True})
    assert abs(res['E'] - g['xxz_E']) < 1e-10 * abs(g['xxz_E'])
    assert np.max(np.abs(psi.entanglement_entropy() - g['xxz_S'])) < 1e-7
