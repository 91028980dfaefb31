"""Truncation of Schmidt values and the truncated SVD of the two-site wave function.

Host-side mirror of the reference ``tenpy/linalg/truncation.py`` (`TruncationError` :57, `truncate` :146,
`svd_theta` :258).  `truncate` works on the 1-D singular value vector on the host (it is tiny), exactly like
the reference; the SVD itself and the compression of `U`, `VH` (``iproject``) run on the device.
"""
# Copyright (C) 2026 tenpy_b200 authors. Apache-2.0.

import warnings

import numpy as np

from . import np_conserved as npc

__all__ = ['TruncationError', 'truncate', 'svd_theta', 'decompose_theta_qr_based']


class TruncationError:
    r"""Accumulated truncation error ``eps`` and overlap bound ``ov`` (reference truncation.py:57)."""

    def __init__(self, eps=0., ov=1.):
        self.eps = eps
        self.ov = ov

    def copy(self):
        return TruncationError(self.eps, self.ov)

    @classmethod
    def from_norm(cls, norm_new, norm_old=1.):
        eps = 1. - norm_new**2 / norm_old**2
        return cls(eps, 1. - 2. * eps)

    @classmethod
    def from_S(cls, S_discarded, norm_old=None):
        eps = np.sum(np.square(S_discarded))
        if norm_old:
            eps /= norm_old * norm_old
        return cls(eps, 1. - 2. * eps)

    def __add__(self, other):
        res = TruncationError()
        res.eps = self.eps + other.eps
        res.ov = self.ov * other.ov
        return res

    @property
    def ov_err(self):
        return 1. - self.ov

    def __repr__(self):
        return 'TruncationError(eps={0:.4e}, ov={1:.10f})'.format(self.eps, self.ov)


def _combine_constraints(good1, good2, warn):
    """Logical and of two constraints, unless that leaves nothing (reference truncation.py:719)."""
    res = np.logical_and(good1, good2)
    if np.any(res):
        return res
    warnings.warn('truncation: can not satisfy constraint for ' + warn, stacklevel=3)
    return good1


def truncate(S, options):
    """Decide which Schmidt values to keep (reference truncation.py:146).

    Options: `chi_max` (100), `chi_min`, `degeneracy_tol`, `svd_min` (1e-14), `trunc_cut` (1e-14).
    Returns ``(mask, norm_new, TruncationError)``."""
    chi_max = options.get('chi_max', 100)
    chi_min = options.get('chi_min', None)
    deg_tol = options.get('degeneracy_tol', None)
    svd_min = options.get('svd_min', 1.e-14)
    trunc_cut = options.get('trunc_cut', 1.e-14)
    if trunc_cut is not None and trunc_cut >= 1.:
        raise ValueError('trunc_cut >=1.')
    S = np.asarray(S)
    if not np.any(S > 1.e-10):
        warnings.warn('no Schmidt value above 1.e-10', stacklevel=2)
    if np.any(S < -1.e-10):
        warnings.warn('negative Schmidt values!', stacklevel=2)
    logS = np.log(np.choose(S <= 0., [S, 1.e-100 * np.ones(len(S))]))
    piv = np.argsort(logS)
    logS = logS[piv]
    good = np.ones(len(piv), dtype=np.bool_)
    if chi_max is not None:
        good2 = np.zeros(len(piv), dtype=np.bool_)
        good2[-int(chi_max):] = True
        good = _combine_constraints(good, good2, 'chi_max')
    if chi_min is not None and chi_min > 1:
        good2 = np.ones(len(piv), dtype=np.bool_)
        good2[-int(chi_min) + 1:] = False
        good = _combine_constraints(good, good2, 'chi_min')
    if deg_tol:
        good2 = np.empty(len(piv), np.bool_)
        good2[0] = True
        good2[1:] = np.greater_equal(logS[1:] - logS[:-1], deg_tol)
        good = _combine_constraints(good, good2, 'degeneracy_tol')
    if svd_min is not None:
        good2 = np.greater_equal(logS, np.log(svd_min))
        good = _combine_constraints(good, good2, 'svd_min')
    if trunc_cut is not None:
        good2 = (np.cumsum(S[piv]**2) > trunc_cut * trunc_cut)
        good = _combine_constraints(good, good2, 'trunc_cut')
    cut = np.nonzero(good)[0][0]
    mask = np.zeros(len(S), dtype=np.bool_)
    np.put(mask, piv[cut:], True)
    norm_new = np.linalg.norm(S[mask])
    return mask, norm_new, TruncationError.from_S(S[np.logical_not(mask)])


subspace_stats = {'tried': 0, 'used': 0, 'residuals': []}   # diagnostics of the subspace warm start


def _subspace_svd(theta, subspace, tol, chi_max, qtotal_LR, inner_labels):
    """SVD of `theta` restricted to the span of the previously kept left (or right) singular vectors.

    With ``P = U_k U_k^dagger`` the part ``(1 - P) theta`` is what a two-site update adds to the old basis; once the
    state has converged its norm is below the truncation tolerance and ``theta`` can be decomposed inside the old
    subspace: ``theta ~= U_k svd(U_k^dagger theta)`` -- a (chi x d chi) instead of a (d chi x d chi) problem whose
    rows are already nearly orthogonal, graded and sorted, and the kept isometry needs no completion.
    Returns ``None`` (caller falls back to the full SVD) unless ``|(1-P) theta| <= tol |theta|``."""
    Uk, VHk = subspace
    chinfo = theta.chinfo
    qtotal_L, qtotal_R = qtotal_LR
    if qtotal_L is None and qtotal_R is None:
        qtotal_R = theta.qtotal
    if qtotal_L is None:
        qtotal_L = chinfo.make_valid(theta.qtotal - qtotal_R)
    elif qtotal_R is None:
        qtotal_R = chinfo.make_valid(theta.qtotal - qtotal_L)
    qtotal_L, qtotal_R = chinfo.make_valid(qtotal_L), chinfo.make_valid(qtotal_R)
    need = min(chi_max if chi_max is not None else min(theta.shape), min(theta.shape))
    nrm = npc.norm(theta)
    if Uk is not None and Uk.rank == 2 and need <= Uk.shape[1] < theta.shape[0]:
        try:
            Uk.legs[0].test_equal(theta.legs[0])
        except ValueError:
            return None
        subspace_stats['tried'] += 1
        a2 = npc.tensordot(Uk.conj(), theta, axes=[0, 0])
        a2.iset_leg_labels([None, theta._labels[1]])
        back = npc.tensordot(Uk, a2, axes=[1, 0])
        back.iset_leg_labels(theta.get_leg_labels())
        rel = npc.norm(theta - back) / nrm
        subspace_stats['residuals'].append(rel)
        if rel > tol:
            return None
        U2, S, VH = npc.svd(a2, qtotal_LR=[chinfo.make_valid(qtotal_L - Uk.qtotal), qtotal_R],
                            inner_labels=inner_labels, deflation_tol=tol, n_keep=chi_max)
        U = npc.tensordot(Uk, U2, axes=[1, 0])
        U.iset_leg_labels([theta._labels[0], inner_labels[0]])
        subspace_stats['used'] += 1
        return U, S, VH, rel * nrm
    if VHk is not None and VHk.rank == 2 and need <= VHk.shape[0] < theta.shape[1]:
        try:
            VHk.legs[1].test_equal(theta.legs[1])
        except ValueError:
            return None
        subspace_stats['tried'] += 1
        a2 = npc.tensordot(theta, VHk.conj(), axes=[1, 1])
        a2.iset_leg_labels([theta._labels[0], None])
        back = npc.tensordot(a2, VHk, axes=[1, 0])
        back.iset_leg_labels(theta.get_leg_labels())
        rel = npc.norm(theta - back) / nrm
        subspace_stats['residuals'].append(rel)
        if rel > tol:
            return None
        U, S, VH2 = npc.svd(a2, qtotal_LR=[qtotal_L, chinfo.make_valid(qtotal_R - VHk.qtotal)],
                            inner_labels=inner_labels, deflation_tol=tol, n_keep=chi_max)
        VH = npc.tensordot(VH2, VHk, axes=[1, 0])
        VH.iset_leg_labels([inner_labels[1], theta._labels[1]])
        subspace_stats['used'] += 1
        return U, S, VH, rel * nrm
    return None


def svd_theta(theta, trunc_par, qtotal_LR=[None, None], inner_labels=['vR', 'vL'], guess=None, full_out=None,
              subspace=None):
    """SVD of the matrix `theta` and truncation (reference truncation.py:258).

    Returns ``(U, S, VH, err, renormalization)`` with ``theta ~= U diag(S * renormalization) VH``.
    Extensions (all optional, the defaults reproduce the reference's behaviour up to the deflation tolerance):
    ``trunc_par['svd_deflation_tol']`` (default 1e-10): singular directions below that fraction of ``|theta|``
    are not iterated to convergence inside the Jacobi SVD -- their values are reported approximately
    (absolute error below the tolerance) and their vectors are an orthonormal completion; the state changes by
    at most that relative amount, the energy to second order in it.
    `subspace` = ``(U_k, VH_k)``, the truncated isometries kept at this bond by the previous update: see
    :func:`_subspace_svd` (used only if the part of `theta` outside their span is below the same tolerance; its
    weight is added to the truncation error).
    `guess` is handed to :func:`npc.svd` (complete orthonormal bases, warm start); if `full_out` is a list the
    untruncated ``(U, VH)`` are appended to it."""
    tol = trunc_par.get('svd_deflation_tol', 1.e-10)
    chi_max = trunc_par.get('chi_max', 100)
    res = _subspace_svd(theta, subspace, tol, chi_max, qtotal_LR, inner_labels) if subspace is not None else None
    lost = 0.
    if res is not None:
        U, S, VH, lost = res
    else:
        U, S, VH = npc.svd(theta, full_matrices=False, compute_uv=True, qtotal_LR=qtotal_LR,
                           inner_labels=inner_labels, guess=guess, deflation_tol=tol,
                           n_keep=(chi_max if full_out is None else None))
    if full_out is not None:
        full_out.append((U.copy(deep=False), VH.copy(deep=False)))
    renormalization = np.sqrt(np.sum(S**2) + lost**2)
    S = S / renormalization
    piv, new_norm, err = truncate(S, trunc_par)
    if lost:
        err = err + TruncationError.from_norm(np.sqrt(max(0., 1. - (lost / renormalization)**2)))
    new_len_S = np.sum(piv, dtype=np.int_)
    if new_len_S * 100 < len(S) and (trunc_par.get('chi_max', 100) is None or
                                     new_len_S != trunc_par.get('chi_max', 100)):
        warnings.warn('Catastrophic reduction in chi: {0:d} -> {1:d}'.format(len(S), int(new_len_S)), stacklevel=2)
    S = S[piv] / new_norm
    renormalization *= new_norm
    U.iproject(piv, axes=1)
    VH.iproject(piv, axes=0)
    return U, S, VH, err, renormalization


# ----------------------------------------------------------------------------------------------------------------------
# QR based truncation (reference truncation.py:370-713)
def _block_vector_norms(arr, norm_axis):
    """per stored block of the 2D Array `arr`: 2-norms over `norm_axis` (host vectors, layout order) -- the reference
    reads ``np.linalg.norm(block, axis=norm_axis)`` from its host blocks (truncation.py:456); here one squared-norm
    launch per block and a D2H copy of the resulting vectors."""
    from .. import backend
    src = arr if norm_axis == 0 else arr.transpose([1, 0])
    lay = src._layout
    lib = backend.get_lib()
    out = []
    for o, (mm, nn) in zip(lay.offsets, lay.shapes):
        mm, nn = int(mm), int(nn)
        buf = backend.empty(nn)
        lib.col_sqnorms(mm, nn, nn, src._buf[int(o):int(o) + mm * nn], buf)
        out.append(np.sqrt(backend.to_host(buf)))
    return out, lay.qdata[:, 1]      # qindex (of `arr`) along the axis that is NOT summed over


def _qr_theta_Y0(old_qtotal_L, old_qtotal_R, old_bond_leg, theta, move_right, expand, min_block_increase):
    """Initial guess `Y0` of the isometry on the expanded bond: the columns (rows) of `theta` with the largest norms in
    every charge block, ``expand`` times the old bond dimension more than the old leg had (reference truncation.py:370).
    Returns an Array with legs ``[(vL.p0), vR]`` (`move_right`) or ``[vL, (p1.vR)]``."""
    assert min_block_increase >= 0
    assert expand is not None and expand != 0
    Y0 = theta.copy(deep=False)
    if move_right:
        Y0.legs[1] = Y0.legs[1].to_LegCharge()
        Y0.ireplace_label('(p1.vR)', 'vR')
        q_axis, norm_axis, lab = 1, 0, 'vR'
    else:
        Y0.legs[0] = Y0.legs[0].to_LegCharge()
        Y0.ireplace_label('(vL.p0)', 'vL')
        q_axis, norm_axis, lab = 0, 1, 'vL'
    # (the reference calls `Y0.gauge_total_charge(...)` here without using the returned copy: no effect)
    v_old = old_bond_leg
    if not v_old.is_blocked():
        v_old = v_old.sort()[1]
    v_new = Y0.get_leg(lab)                   # blocked: created from a pipe
    piv = np.zeros(v_new.ind_len, dtype=bool)
    increase_per_block = max(min_block_increase, int(v_old.ind_len * expand // v_new.block_number))
    sizes_old = v_old.get_block_sizes()
    sizes_new = v_new.get_block_sizes()
    norms, qidx = _block_vector_norms(Y0, norm_axis)
    by_q = {int(q): nv for q, nv in zip(qidx, norms)}
    j_old = 0
    q_old = v_old.charges[j_old, :]
    for j_new, q_new in enumerate(v_new.charges):
        if np.all(q_new == q_old):            # charge block both in v_new and v_old
            s_new = sizes_old[j_old] + increase_per_block
            j_old += 1
            if j_old < len(v_old.charges):
                q_old = v_old.charges[j_old, :]
            else:
                q_old = None
        else:
            s_new = increase_per_block
        s_new = min(int(s_new), int(sizes_new[j_new]))
        nv = by_q.get(j_new)
        if nv is None:                        # block not stored in theta
            continue
        kept = np.argsort(-nv, kind='stable')[:s_new]
        piv[v_new.slices[j_new] + kept] = True
    Y0.iproject(piv, lab)
    return Y0


def _eig_based_svd(A, need_U=True, need_Vd=True, inner_labels=[None, None], trunc_params=None):
    """Singular values / one set of singular vectors of `A` from the eigen-decomposition of ``A A^dagger`` or
    ``A^dagger A`` (reference truncation.py:473): two GEMM-class contractions and a batched `eigh` instead of an SVD."""
    assert A.rank == 2
    if need_U and need_Vd:
        raise NotImplementedError('both U and Vd from eigh: relative phases are not fixed (as in the reference)')
    U = Vd = None
    if need_U:
        L, U = npc.eigh(npc.tensordot(A, A.conj(), axes=[1, 1]), sort='>')
        U.ireplace_label('eig', inner_labels[0])
    elif need_Vd:
        L, V = npc.eigh(npc.tensordot(A.conj(), A, axes=[0, 0]), sort='>')
        Vd = V.iconj().itranspose().ireplace_label('eig*', inner_labels[1])
    else:
        A2 = npc.tensordot(A, A.conj(), axes=[1, 1]) if A.shape[1] >= A.shape[0] else \
            npc.tensordot(A.conj(), A, axes=[0, 0])
        L = npc.eigvalsh(A2)
    S = np.sqrt(np.abs(L))
    if trunc_params is not None:
        piv, renormalize, trunc_err = truncate(S, trunc_params)
        S = S[piv] / renormalize
        if need_U:
            U.iproject(piv, 1)
        if need_Vd:
            Vd.iproject(piv, 0)
    else:
        renormalize = np.linalg.norm(S)
        S = S / renormalize
        trunc_err = TruncationError()
    return U, S, Vd, trunc_err, renormalize


def decompose_theta_qr_based(old_qtotal_L, old_qtotal_R, old_bond_leg, theta, move_right, expand, min_block_increase,
                             use_eig_based_svd, trunc_params, compute_err, return_both_T):
    """QR based decomposition and truncation of the two-site wave function ``theta[(vL.p0), (p1.vR)]``
    (reference truncation.py:533): two QR steps on an expanded bond (controlled bond expansion) reduce `theta` to a
    small bond matrix ``Xi``, only ``Xi`` is decomposed by an SVD (or `eigh`).

    Returns ``(T_Lc, S, T_Rc, form, trunc_err, renormalization)`` as the reference."""
    if compute_err:
        return_both_T = True
    Y0 = _qr_theta_Y0(old_qtotal_L, old_qtotal_R, old_bond_leg, theta, move_right, expand, min_block_increase)
    if move_right:
        theta_i1 = npc.tensordot(Y0.conj(), theta, axes=['(vL*.p0*)', '(vL.p0)']).ireplace_label('vR*', 'vL')
        theta_i1.itranspose(['(p1.vR)', 'vL'])
        B_R, _ = npc.qr(theta_i1, inner_labels=['vL', 'vR'], inner_qconj=-1)
        B_R.itranspose(['vL', '(p1.vR)'])
        theta_i0 = npc.tensordot(theta, B_R.conj(), axes=['(p1.vR)', '(p1*.vR*)']).ireplace_label('vL*', 'vR')
        A_L, Xi = npc.qr(theta_i0, inner_labels=['vR', 'vL'])
    else:
        theta_i0 = npc.tensordot(theta, Y0.conj(), axes=['(p1.vR)', '(p1*.vR*)']).ireplace_label('vL*', 'vR')
        A_L, _ = npc.qr(theta_i0, inner_labels=['vR', 'vL'])
        theta_i1 = npc.tensordot(A_L.conj(), theta, axes=['(vL*.p0*)', '(vL.p0)']).ireplace_label('vR*', 'vL')
        theta_i1.itranspose(['(p1.vR)', 'vL'])
        B_R, Xi = npc.qr(theta_i1, inner_labels=['vL', 'vR'], inner_qconj=-1)
        B_R.itranspose(['vL', '(p1.vR)'])
        Xi.itranspose(['vL', 'vR'])
    if use_eig_based_svd:
        U, S, Vd, _, renormalization = _eig_based_svd(Xi, need_U=move_right, need_Vd=(not move_right),
                                                      inner_labels=['vR', 'vL'], trunc_params=trunc_params)
    else:
        U, S, Vd, _, renormalization = svd_theta(Xi, trunc_params)
    T_Lc, T_Rc = None, None
    form = ['A', 'B']
    if move_right:
        T_Lc = npc.tensordot(A_L, U, axes=['vR', 'vL'])
        if return_both_T:
            if use_eig_based_svd:
                T_Rc = npc.tensordot(Xi, B_R, axes=['vR', 'vL'])
                T_Rc = npc.tensordot(U.conj(), T_Rc, axes=['vL*', 'vL']).ireplace_label('vR*', 'vL')
                T_Rc = T_Rc / npc.norm(T_Rc)
                form[1] = 'Th'
            else:
                T_Rc = npc.tensordot(Vd, B_R, axes=['vR', 'vL'])
    else:
        T_Rc = npc.tensordot(Vd, B_R, axes=['vR', 'vL'])
        if return_both_T:
            if use_eig_based_svd:
                T_Lc = npc.tensordot(A_L, Xi, axes=['vR', 'vL'])
                T_Lc = npc.tensordot(T_Lc, Vd.conj(), axes=['vR', 'vR*']).ireplace_label('vL*', 'vR')
                T_Lc = T_Lc / npc.norm(T_Lc)
                form[0] = 'Th'
            else:
                T_Lc = npc.tensordot(A_L, U, axes=['vR', 'vL'])
    if compute_err:
        if use_eig_based_svd:
            theta_approx = npc.tensordot(T_Lc, T_Rc, axes=['vR', 'vL'])
        else:
            theta_approx = npc.tensordot(T_Lc.scale_axis(S, axis='vR'), T_Rc, axes=['vR', 'vL'])
        N_theta = npc.norm(theta)
        eps = npc.norm(theta / N_theta - theta_approx * (renormalization / N_theta))**2
        trunc_err = TruncationError(eps, 1. - 2. * eps)
    else:
        trunc_err = TruncationError(np.nan, np.nan)
    if T_Lc is not None:
        T_Lc.ireplace_label('(vL.p0)', '(vL.p)')
    if T_Rc is not None:
        T_Rc.ireplace_label('(p1.vR)', '(p.vR)')
    return T_Lc, S, T_Rc, form, trunc_err, renormalization
