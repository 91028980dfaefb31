"""CPU tests pinning the ORACLE (oracle/*.py) against the golden vectors generated from the reference.

Integer results (block tables, charges, slices, q_map, truncation masks) must match exactly; floating point
blocks within 1e-13 (same LAPACK/BLAS underneath, different call order); singular values / eigenvalues are
compared sorted, as the reference's tests do (tests/test_np_conserved.py:674)."""
import numpy as np
import pytest

import helpers as h
from oracle import npc_blocks as ob
from oracle import dmrg_dense as od


@pytest.fixture(scope='module')
def g_td():
    return h.load('tensordot.npz')


@pytest.fixture(scope='module')
def g_rs():
    return h.load('reshape_svd.npz')


@pytest.fixture(scope='module')
def g_dm():
    return h.load('dmrg.npz')


def test_tensordot_inner_norm_add(g_td):
    for ci in range(int(g_td['ncases'])):
        a, b, c = (h.oarray_from(g_td, 'c%d_%s' % (ci, k)) for k in 'abc')
        n = int(g_td['c%d_naxes' % ci])
        if n == 2:
            res = ob.tensordot(a, b, 2)
        else:
            res = ob.tensordot(a, b, 1)
        h.assert_close(res, c)
        assert np.array_equal(res.qdata, c.qdata), 'result must be lex-sorted like the reference (pyx:1777)'
        assert abs(ob.inner(a, a, True) - g_td['c%d_inner_aa' % ci]) < 1e-12
        assert abs(ob.norm(a) - g_td['c%d_norm_a' % ci]) < 1e-13
        a2 = h.oarray_from(g_td, 'c%d_a2' % ci)
        assert abs(ob.inner(a, a2, True) - g_td['c%d_inner_aa2' % ci]) < 1e-12
        h.assert_close(ob.iadd_prefactor_other(a, 0.37, a2), h.oarray_from(g_td, 'c%d_sum' % ci))


def test_combine_split_transpose(g_rs):
    a = h.oarray_from(g_rs, 'a')
    comb = h.oarray_from(g_rs, 'comb')
    res = ob.combine_legs(a, [[0, 1], [3, 4]], [0, 2], [comb.legs[0], comb.legs[2]])
    h.assert_close(res, comb, 0.)
    comb2 = h.oarray_from(g_rs, 'comb2')
    res2 = ob.combine_legs(a, [[4, 3], [1, 0]], [0, 2], [comb2.legs[0], comb2.legs[2]])
    h.assert_close(res2, comb2, 0.)
    h.assert_close(ob.split_legs(comb, [0, 2]), h.oarray_from(g_rs, 'split'), 0.)
    h.assert_close(ob.transpose(a, [3, 0, 2, 4, 1]), h.oarray_from(g_rs, 'transp'), 0.)


def test_svd_eigh_project_scale(g_rs):
    m = h.oarray_from(g_rs, 'm')
    U, S, VH = ob.svd(m)
    Uref = h.oarray_from(g_rs, 'm_U')
    h.assert_same_structure(U, Uref)
    assert np.max(np.abs(S - g_rs['m_S'])) < 1e-13            # block-ordered, descending inside a block
    rec = ob.tensordot(ob.scale_axis(U, S, 1), VH, 1)
    h.assert_close(rec, m, 1e-13, structure=False)
    U2, S2, _ = ob.svd(m, qtotal_LR=([1, 1], None), inner_qconj=-1)
    h.assert_same_structure(U2, h.oarray_from(g_rs, 'm_U2'))
    # svd_theta = svd + truncate + project
    Sn = S / np.linalg.norm(S)
    mask, new_norm, err = od.truncate(Sn, chi_max=17, svd_min=1e-8)
    assert np.max(np.abs(Sn[mask] / new_norm - g_rs['m_St'])) < 1e-13
    assert abs(err - g_rs['m_err']) < 1e-15
    h.assert_same_structure(ob.project(U, mask, 1), h.oarray_from(g_rs, 'm_Ut'))
    h.assert_close(ob.project(m, g_rs['proj_mask'], 1), h.oarray_from(g_rs, 'm_proj'), 0.)
    h.assert_close(ob.scale_axis(m, g_rs['scale_s'], 0), h.oarray_from(g_rs, 'm_scaled'), 1e-15)
    rho = h.oarray_from(g_rs, 'rho')
    w, V = ob.eigh(rho)
    assert np.max(np.abs(w - g_rs['rho_w'])) < 1e-12
    h.assert_same_structure(V, h.oarray_from(g_rs, 'rho_V'))
    Vd = V.to_dense()
    assert np.max(np.abs(Vd @ np.diag(w) @ Vd.T - rho.to_dense())) < 1e-12


def test_truncate(g_rs):
    S = g_rs['trunc_S']
    opts = [dict(chi_max=10), dict(chi_max=30, svd_min=1e-4), dict(chi_max=100, trunc_cut=1e-3),
            dict(chi_max=12, chi_min=5, degeneracy_tol=1e-2)]
    for k, o in enumerate(opts):
        mask, nn, err = od.truncate(S, **o)
        assert np.array_equal(mask, g_rs['trunc%d_mask' % k])
        assert abs(nn - g_rs['trunc%d_norm' % k]) < 1e-15
        assert abs(err - g_rs['trunc%d_err' % k]) < 1e-15


def test_two_site_matvec_and_env(g_dm):
    LHeff, RHeff, theta = (h.oarray_from(g_dm, 'xxz_' + k) for k in ('LHeff', 'RHeff', 'theta'))
    Hth = ob.two_site_matvec(LHeff, RHeff, theta)
    ref = h.oarray_from(g_dm, 'xxz_Htheta')
    # the reference transposes the result back to theta's leg order; labels (vL.p0), (p1.vR)
    h.assert_close(Hth, ref, 1e-13)
    # the energy <theta|H|theta> of the converged state equals the DMRG energy
    E = ob.inner(theta, Hth, True) / ob.inner(theta, theta, True)
    assert abs(E - g_dm['xxz_E']) < 1e-10
    U = h.oarray_from(g_dm, 'xxz_U')
    LP = ob.tensordot(LHeff, U, 1)
    Uc = ob.OArray([l.conj() for l in U.legs], U.mod, -U.qtotal, U.qdata, U.blocks)
    LP = ob.tensordot(ob.transpose(Uc, [1, 0]), LP, 1)
    h.assert_close(LP, h.oarray_from(g_dm, 'xxz_LPnew'), 1e-12)
    _, S, _ = ob.svd(theta)
    # golden S belongs to the Lanczos-polished theta (may contain a few more tiny blocks): compare the top values
    assert np.max(np.abs(np.sort(S)[::-1][:40] - np.sort(g_dm['xxz_theta_S'])[::-1][:40])) < 1e-8


def test_dense_dmrg_tfi(g_dm):
    """oracle dense DMRG reproduces BASELINE.md config 1: E = -25.1077971116238 (reference, same options)"""
    r = od.run_dmrg(od.tfi_mpo(1., 1.), 20, 2, [0] * 20, dict(chi_max=50, svd_min=1e-10), {}, max_E_err=1e-10)
    assert abs(r['E'] - g_dm['tfi_E']) < 1e-10 * abs(g_dm['tfi_E'])
    assert abs(g_dm['tfi_E'] - (-25.1077971116238)) < 1e-12
    assert np.max(np.abs(np.array(r['S']) - g_dm['tfi_S'])) < 1e-8
    sv = r['Ss'][10]
    ref = g_dm['tfi_sv_mid']
    k = min(len(sv), len(ref))
    assert np.max(np.abs(np.sort(sv)[::-1][:k] - np.sort(ref)[::-1][:k])) < 1e-8


def test_lanczos_dense():
    """oracle Lanczos vs dense eigh (reference tests/test_krylov_based.py:33, tolerance 5e-14 there)"""
    rng = np.random.default_rng(7)
    for n in (4, 20, 60):
        Hm = rng.standard_normal((n, n))
        Hm = Hm + Hm.T
        E0, psi, N = od.lanczos_ground(lambda x: Hm @ x, rng.standard_normal(n), N_max=n + 2, P_tol=1e-28)
        w, v = np.linalg.eigh(Hm)
        assert abs(E0 - w[0]) < 1e-10
        assert abs(abs(np.dot(psi, v[:, 0])) - 1.) < 1e-8


def test_dense_matvec_orders_agree():
    """the two contraction orders of TwoSiteH.matvec in the reference (combine=True :1337, combine=False :1340)
    are the same linear map; pins `od.matvec_split` on `od.matvec` (itself pinned on the golden vectors above)"""
    rng = np.random.default_rng(5)
    chi_l, chi_r, d, D = 7, 5, 2, 3
    LP = rng.standard_normal((chi_l, D, chi_l))
    RP = rng.standard_normal((chi_r, D, chi_r))
    W0 = od.tfi_mpo(1.3, 0.7)
    W1 = rng.standard_normal((D, D, d, d))
    theta = rng.standard_normal((chi_l, d, d, chi_r))
    a = od.matvec(od.contract_LHeff(LP, W0), od.contract_RHeff(RP, W1), theta.reshape(chi_l * d, d * chi_r))
    b = od.matvec_split(LP, W0, W1, RP, theta).reshape(chi_l * d, d * chi_r)
    assert np.max(np.abs(a - b)) < 1e-13 * np.max(np.abs(a))
