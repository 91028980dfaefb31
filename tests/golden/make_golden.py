#!/usr/bin/env python
"""Generate the golden vectors of tests/golden/*.npz by running the UNMODIFIED reference (tenpy/tenpy,
pure-Python path) on seeded inputs.  Run in the build container only (needs /root/reference):

    TENPY_NO_CYTHON=1 PYTHONPATH=/root/reference python tests/golden/make_golden.py

The reference cannot travel to the GPU box, so the outputs are committed as small fixtures.  Each file
stores inputs and the reference's outputs in a flat dict of numpy arrays (see `dump_array`).
"""
import os
import sys
import warnings

import numpy as np

os.environ.setdefault('TENPY_NO_CYTHON', '1')
sys.path.insert(0, os.environ.get('TENPY_REFERENCE', '/root/reference'))
warnings.simplefilter('ignore')

import tenpy  # noqa: E402
import tenpy.linalg.np_conserved as npc  # noqa: E402
from tenpy.linalg import charges as rcharges  # noqa: E402
from tenpy.linalg.truncation import truncate, svd_theta  # noqa: E402
from tenpy.linalg.krylov_based import LanczosGroundState  # noqa: E402
from tenpy.algorithms import dmrg  # noqa: E402
from tenpy.algorithms.mps_common import TwoSiteH  # noqa: E402
from tenpy.networks.mps import MPS  # noqa: E402
from tenpy.models.tf_ising import TFIChain  # noqa: E402
from tenpy.models.spins import SpinChain  # noqa: E402
from tenpy.models.hubbard import FermiHubbardChain  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def dump_leg(prefix, leg, out):
    out[prefix + '_slices'] = np.asarray(leg.slices, dtype=np.int64)
    out[prefix + '_charges'] = np.asarray(leg.charges, dtype=np.int64)
    out[prefix + '_qconj'] = np.int64(leg.qconj)
    if isinstance(leg, rcharges.LegPipe):
        out[prefix + '_pipe_nlegs'] = np.int64(leg.nlegs)
        out[prefix + '_pipe_qmap'] = np.asarray(leg.q_map, dtype=np.int64)
        out[prefix + '_pipe_qmap_slices'] = np.asarray(leg.q_map_slices, dtype=np.int64)
        for j, sub in enumerate(leg.legs):
            dump_leg(prefix + '_sub%d' % j, sub, out)


def dump_array(prefix, arr, out):
    arr = arr.copy(deep=True)
    arr.isort_qdata()
    arr._imake_contiguous()
    out[prefix + '_nlegs'] = np.int64(arr.rank)
    out[prefix + '_mod'] = np.asarray(arr.chinfo.mod, dtype=np.int64)
    out[prefix + '_qtotal'] = np.asarray(arr.qtotal, dtype=np.int64)
    out[prefix + '_qdata'] = np.asarray(arr._qdata, dtype=np.int64).reshape(-1, arr.rank)
    out[prefix + '_data'] = np.concatenate([np.asarray(b, dtype=np.float64).reshape(-1) for b in arr._data]) \
        if arr.stored_blocks else np.zeros(0)
    out[prefix + '_labels'] = np.array([l if l is not None else '' for l in arr.get_leg_labels()])
    for i, leg in enumerate(arr.legs):
        dump_leg(prefix + '_leg%d' % i, leg, out)


def rand_leg(rng, chinfo, n, qconj, spread=2):
    q = chinfo.make_valid(rng.integers(-spread, spread + 1, size=(n, chinfo.qnumber)))
    if rng.random() < 0.7 and chinfo.qnumber > 0:
        q = q[np.lexsort(q.T)]
    return npc.LegCharge.from_qflat(chinfo, q, qconj)


def rand_array(rng, legs, qtotal=None, labels=None):
    return npc.Array.from_func(rng.standard_normal, legs, qtotal=qtotal, labels=labels)


def golden_tensordot():
    out = {}
    rng = np.random.default_rng(3141592)
    cases = []
    ch1 = npc.ChargeInfo([1], ['U1'])
    ch2 = npc.ChargeInfo([1, 2], ['N', 'P'])
    ch3 = npc.ChargeInfo([3], ['Z3'])
    ch0 = npc.ChargeInfo()
    for ci, (ch, naxes) in enumerate([(ch1, 1), (ch1, 2), (ch2, 2), (ch3, 1), (ch0, 2), (ch2, 1)]):
        la, lb, lc, ld = rand_leg(rng, ch, 12, +1), rand_leg(rng, ch, 7, -1), rand_leg(rng, ch, 9, +1), \
            rand_leg(rng, ch, 6, -1)
        qa = ch.make_valid(rng.integers(-1, 2, size=ch.qnumber))
        qb = ch.make_valid(rng.integers(-1, 2, size=ch.qnumber))
        a = rand_array(rng, [la, lb, lc], qa)
        if naxes == 2:
            b = rand_array(rng, [lb.conj(), lc.conj(), ld], qb)
        else:
            b = rand_array(rng, [lc.conj(), ld, la.conj()], qb)
        c = npc.tensordot(a, b, axes=naxes)
        dump_array('c%d_a' % ci, a, out)
        dump_array('c%d_b' % ci, b, out)
        dump_array('c%d_c' % ci, c, out)
        out['c%d_naxes' % ci] = np.int64(naxes)
        out['c%d_inner_aa' % ci] = np.float64(npc.inner(a, a, axes='range', do_conj=True))
        out['c%d_norm_a' % ci] = np.float64(npc.norm(a))
        a2 = rand_array(rng, [la, lb, lc], qa)
        # drop some blocks of a2 so that the block tables differ
        a2d = a2.to_ndarray()
        sl = la.get_slice(0)
        a2d[sl] = 0.
        a2 = npc.Array.from_ndarray(a2d, [la, lb, lc], qtotal=qa, cutoff=1e-14)
        dump_array('c%d_a2' % ci, a2, out)
        dump_array('c%d_sum' % ci, a + a2 * 0.37, out)
        out['c%d_inner_aa2' % ci] = np.float64(npc.inner(a, a2, axes='range', do_conj=True))
        cases.append(ci)
    out['ncases'] = np.int64(len(cases))
    np.savez_compressed(os.path.join(HERE, 'tensordot.npz'), **out)


def golden_reshape_svd():
    out = {}
    rng = np.random.default_rng(2718281)
    ch = npc.ChargeInfo([1, 1], ['N', 'Sz'])
    lv = rand_leg(rng, ch, 14, +1, spread=2)
    lp = npc.LegCharge.from_qflat(ch, [[0, 0], [1, 1], [1, -1], [2, 0]], +1)
    lw = npc.LegCharge.from_qind(ch, np.arange(5), [[0, 0], [1, 1], [-1, -1], [0, 0]], -1)
    lr = rand_leg(rng, ch, 11, -1, spread=2)
    a = rand_array(rng, [lv, lp, lw, lp.conj(), lr], [1, 1], ['vL', 'p0', 'w', 'p1', 'vR'])
    dump_array('a', a, out)
    comb = a.combine_legs([['vL', 'p0'], ['p1', 'vR']], qconj=[+1, -1])
    dump_array('comb', comb, out)
    comb2 = a.combine_legs([['vR', 'p1'], ['p0', 'vL']], new_axes=[0, 2], qconj=[-1, +1])
    dump_array('comb2', comb2, out)
    dump_array('split', comb.split_legs(), out)
    dump_array('transp', a.transpose(['p1', 'vL', 'w', 'vR', 'p0']), out)
    # a matrix for svd / eigh / project / scale_axis
    m = rand_array(rng, [lv, lp, lp.conj(), lr], [0, 0], ['vL', 'p0', 'p1', 'vR']).combine_legs(
        [['vL', 'p0'], ['p1', 'vR']], qconj=[+1, -1])
    dump_array('m', m, out)
    U, S, VH = npc.svd(m, inner_labels=['vR', 'vL'])
    dump_array('m_U', U, out)
    dump_array('m_VH', VH, out)
    out['m_S'] = S
    U2, S2, VH2 = npc.svd(m, qtotal_LR=[[1, 1], None], inner_qconj=-1)
    dump_array('m_U2', U2, out)
    out['m_S2'] = S2
    trunc = {'chi_max': 17, 'svd_min': 1e-8}
    Ut, St, VHt, err, renorm = svd_theta(m, trunc, inner_labels=['vR', 'vL'])
    out['m_St'] = St
    out['m_err'] = np.float64(err.eps)
    out['m_renorm'] = np.float64(renorm)
    dump_array('m_Ut', Ut, out)
    mask = rng.random(m.shape[1]) < 0.6
    out['proj_mask'] = mask
    mp = m.copy()
    mp.iproject(mask, 1)
    dump_array('m_proj', mp, out)
    s = rng.random(m.shape[0])
    out['scale_s'] = s
    dump_array('m_scaled', m.scale_axis(s, 0), out)
    rho = npc.tensordot(m, m.conj(), axes=[1, 1])
    dump_array('rho', rho, out)
    w, V = npc.eigh(rho)
    out['rho_w'] = w
    dump_array('rho_V', V, out)
    # truncate
    Sv = np.sort(rng.random(40))[::-1] ** 6
    Sv /= np.linalg.norm(Sv)
    out['trunc_S'] = Sv
    for k, opt in enumerate([{'chi_max': 10}, {'chi_max': 30, 'svd_min': 1e-4}, {'chi_max': 100, 'trunc_cut': 1e-3},
                             {'chi_max': 12, 'chi_min': 5, 'degeneracy_tol': 1e-2}]):
        mask_t, nn, er = truncate(Sv, dict(opt))
        out['trunc%d_mask' % k] = mask_t
        out['trunc%d_norm' % k] = np.float64(nn)
        out['trunc%d_err' % k] = np.float64(er.eps)
    np.savez_compressed(os.path.join(HERE, 'reshape_svd.npz'), **out)


def run_ref_dmrg(model, psi, params):
    eng = dmrg.TwoSiteDMRGEngine(psi, model, params)
    E, _ = eng.run()
    return eng, E


def golden_dmrg():
    out = {}
    # config 1 of BASELINE.json: TFIChain L=20 chi=50
    M = TFIChain(dict(L=20, J=1., g=1., bc_MPS='finite', conserve=None))
    psi = MPS.from_product_state(M.lat.mps_sites(), ['up'] * 20, bc='finite')
    eng, E = run_ref_dmrg(M, psi, {'mixer': None, 'max_E_err': 1e-10, 'trunc_params': {'chi_max': 50, 'svd_min': 1e-10},
                                   'combine': True})
    out['tfi_E'] = np.float64(E)
    out['tfi_S'] = psi.entanglement_entropy()
    out['tfi_chi'] = np.array(psi.chi)
    out['tfi_sv_mid'] = psi.get_SL(10)
    out['tfi_sweeps'] = np.int64(eng.sweeps)
    # eff. H at the centre: matvec / Lanczos golden on an uncharged problem
    # XXZ with Sz conservation + mixer (block-sparse path), small version of config 3
    L = 16
    M = SpinChain(dict(L=L, S=0.5, Jx=1., Jy=1., Jz=1., bc_MPS='finite', conserve='Sz'))
    psi = MPS.from_product_state(M.lat.mps_sites(), ['up', 'down'] * (L // 2), bc='finite')
    params = {'mixer': True, 'mixer_params': {'amplitude': 1e-5, 'decay': 2., 'disable_after': 6},
              'max_E_err': 1e-11, 'max_S_err': 1e-8, 'trunc_params': {'chi_max': 60, 'svd_min': 1e-10},
              'combine': True, 'max_sweeps': 20}
    eng, E = run_ref_dmrg(M, psi, params)
    out['xxz_E'] = np.float64(E)
    out['xxz_S'] = psi.entanglement_entropy()
    out['xxz_chi'] = np.array(psi.chi)
    out['xxz_sv_mid'] = np.sort(psi.get_SL(L // 2))[::-1]
    out['xxz_sweeps'] = np.int64(eng.sweeps)
    # hot-path tensors of the converged state at the centre bond
    i0 = L // 2 - 1
    H = TwoSiteH(eng.env, i0, combine=True)
    theta = H.combine_theta(psi.get_theta(i0, 2))
    dump_array('xxz_LHeff', H.LHeff, out)
    dump_array('xxz_RHeff', H.RHeff, out)
    dump_array('xxz_theta', theta, out)
    dump_array('xxz_Htheta', H.matvec(theta), out)
    E0, th0, N = LanczosGroundState(H, theta, {}).run()
    out['xxz_lanczos_E0'] = np.float64(E0)
    out['xxz_lanczos_N'] = np.int64(N)
    U, S, VH = npc.svd(th0, inner_labels=['vR', 'vL'])
    out['xxz_theta_S'] = S
    LP = npc.tensordot(H.LHeff, U.replace_label('(vL.p0)', '(vL.p)'), axes=['(vR.p0*)', '(vL.p)'])
    LP = npc.tensordot(U.replace_label('(vL.p0)', '(vL.p)').conj(), LP, axes=['(vL*.p*)', '(vR*.p0)'])
    dump_array('xxz_U', U, out)
    dump_array('xxz_LPnew', LP, out)
    # Hubbard with (N, Sz): L=6 is exact at chi=64
    L = 6
    M = FermiHubbardChain(dict(L=L, t=1., U=4., mu=0., bc_MPS='finite', cons_N='N', cons_Sz='Sz'))
    psi = MPS.from_product_state(M.lat.mps_sites(), ['up', 'down'] * (L // 2), bc='finite')
    params = {'mixer': True, 'mixer_params': {'amplitude': 1e-5, 'decay': 2., 'disable_after': 6},
              'max_E_err': 1e-11, 'max_S_err': 1e-8, 'trunc_params': {'chi_max': 64, 'svd_min': 1e-10},
              'combine': True, 'max_sweeps': 20}
    eng, E = run_ref_dmrg(M, psi, params)
    out['hub_E'] = np.float64(E)
    out['hub_S'] = psi.entanglement_entropy()
    out['hub_chi'] = np.array(psi.chi)
    out['hub_sv_mid'] = np.sort(psi.get_SL(L // 2))[::-1]
    dump_array('hub_W', M.H_MPO.get_W(2), out)
    # TFI with parity conservation
    M = TFIChain(dict(L=12, J=1., g=0.8, bc_MPS='finite', conserve='parity'))
    psi = MPS.from_product_state(M.lat.mps_sites(), ['up'] * 12, bc='finite')
    eng, E = run_ref_dmrg(M, psi, {'mixer': True, 'mixer_params': {'disable_after': 5}, 'max_E_err': 1e-11,
                                   'trunc_params': {'chi_max': 40, 'svd_min': 1e-10}, 'combine': True,
                                   'max_sweeps': 16})
    out['tfip_E'] = np.float64(E)
    out['tfip_S'] = psi.entanglement_entropy()
    np.savez_compressed(os.path.join(HERE, 'dmrg.npz'), **out)
    print('tfi E', out['tfi_E'], 'xxz E', out['xxz_E'], 'hub E', out['hub_E'], 'tfip E', out['tfip_E'])


if __name__ == '__main__':
    print('reference tenpy', tenpy.__version__, 'at', os.path.dirname(tenpy.__file__))
    golden_tensordot()
    golden_reshape_svd()
    golden_dmrg()
    for f in sorted(os.listdir(HERE)):
        if f.endswith('.npz'):
            print(f, os.path.getsize(os.path.join(HERE, f)), 'bytes')
