"""Truncation of Schmidt values and the truncated SVD of the two-site wave function.

Host-side mirror of the reference ``tenpy/linalg/truncation.py`` (`TruncationError` :57, `truncate` :146,
`svd_theta` :258).  `truncate` works on the 1-D singular value vector on the host (it is tiny), exactly like
the reference; the SVD itself and the compression of `U`, `VH` (``iproject``) run on the device.
"""
# Copyright (C) 2026 tenpy_b200 authors. Apache-2.0.

import warnings

import numpy as np

from . import np_conserved as npc

__all__ = ['TruncationError', 'truncate', 'svd_theta']


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
