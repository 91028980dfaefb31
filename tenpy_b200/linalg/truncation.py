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


class _KeepCounts:
    """The admissible numbers of kept Schmidt values, narrowed rule by rule.  A rule that would leave no admissible count is
    skipped with a warning (behaviour of the reference's constraint chain, truncation.py:719-731)."""

    def __init__(self, n):
        self.ok = np.ones(n + 1, dtype=np.bool_)     # ok[k]: keeping the k largest values is admissible
        self.ok[0] = False                            # at least one value is always kept

    def require(self, allowed, rule):
        both = self.ok & allowed
        if both.any():
            self.ok = both
        else:
            warnings.warn('truncation: can not satisfy constraint for ' + rule, stacklevel=4)

    def largest(self):
        return int(np.nonzero(self.ok)[0][-1])


def truncate(S, options):
    """Which Schmidt values survive a truncation (reference truncation.py:146; same options, defaults and results).

    The values are ranked in descending order and every option turns into a condition on the NUMBER k of kept values:
    ``chi_max``: k <= chi_max; ``chi_min``: k >= chi_min; ``degeneracy_tol``: no cut between two values whose logarithms are
    closer than the tolerance; ``svd_min``: only values >= svd_min; ``trunc_cut``: the discarded weight
    ``sum_{i >= k} S_i^2`` stays <= trunc_cut^2.  The largest admissible k wins; conditions are imposed in this order and a
    condition that contradicts the earlier ones is dropped with a warning.  Returns ``(mask, norm_new, TruncationError)``."""
    chi_max, chi_min = options.get('chi_max', 100), options.get('chi_min', None)
    deg_tol, svd_min, trunc_cut = options.get('degeneracy_tol', None), options.get('svd_min', 1.e-14), \
        options.get('trunc_cut', 1.e-14)
    if trunc_cut is not None and trunc_cut >= 1.:
        raise ValueError('trunc_cut >=1.')
    S = np.asarray(S)
    n = len(S)
    if not (S > 1.e-10).any():
        warnings.warn('no Schmidt value above 1.e-10', stacklevel=2)
    if (S < -1.e-10).any():
        warnings.warn('negative Schmidt values!', stacklevel=2)
    logs = np.log(np.where(S > 0., S, 1.e-100))
    rank = np.argsort(logs, kind='stable')[::-1]         # rank[0] = position of the largest value
    logs_desc = logs[rank]
    k = np.arange(n + 1)
    counts = _KeepCounts(n)
    if chi_max is not None:
        counts.require(k <= int(chi_max), 'chi_max')
    if chi_min is not None and chi_min > 1:
        counts.require(k >= int(chi_min), 'chi_min')
    if deg_tol:
        gap_ok = np.ones(n + 1, dtype=np.bool_)           # cutting after the k-th value: needs a log gap to the (k+1)-th
        gap_ok[1:n] = (logs_desc[:-1] - logs_desc[1:]) >= deg_tol
        counts.require(gap_ok, 'degeneracy_tol')
    if svd_min is not None:
        counts.require(k <= int(np.count_nonzero(logs_desc >= np.log(svd_min))), 'svd_min')
    if trunc_cut is not None:
        tail = np.concatenate([np.cumsum((S[rank][::-1])**2)[::-1], [0.]])    # tail[k] = weight discarded when keeping k
        enough = np.zeros(n + 1, dtype=np.bool_)
        enough[1:] = tail[:-1] > trunc_cut * trunc_cut       # dropping the k-th value as well would exceed the budget
        counts.require(enough, 'trunc_cut')
    keep = counts.largest()
    mask = np.zeros(n, dtype=np.bool_)
    mask[rank[:keep]] = True
    return mask, np.linalg.norm(S[mask]), TruncationError.from_S(S[~mask])


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
    ``trunc_par['svd_deflation_tol']``: singular directions below that fraction of ``|theta|`` are not iterated to
    convergence inside the Jacobi SVD -- their values are reported approximately (absolute error below the tolerance) and
    their vectors are an orthonormal completion; the state changes by at most that relative amount, the energy to second
    order in it.  Default: ``min(1e-10, svd_min)`` -- never above the smallest Schmidt value the truncation may keep, so
    that every KEPT value is a converged singular value (with ``svd_min=None`` or 0: the rounding-level deflation of
    `npc.svd` only).  A caller that keeps values below 1e-10 on purpose (the benchmark harness: ``svd_min=1e-45`` to hold
    chi saturated) passes the tolerance explicitly.
    `subspace` = ``(U_k, VH_k)``, the truncated isometries kept at this bond by the previous update: see
    :func:`_subspace_svd` (used only if the part of `theta` outside their span is below the same tolerance; its
    weight is added to the truncation error).
    `guess` is handed to :func:`npc.svd` (complete orthonormal bases, warm start); if `full_out` is a list the
    untruncated ``(U, VH)`` are appended to it."""
    tol = trunc_par.get('svd_deflation_tol', None)
    if tol is None:
        svd_min = trunc_par.get('svd_min', 1.e-14)
        tol = min(1.e-10, svd_min) if svd_min else 0.
    chi_max = trunc_par.get('chi_max', 100)
    res = _subspace_svd(theta, subspace, tol, chi_max, qtotal_LR, inner_labels) if subspace is not None else None
    lost = 0.
    if res is not None:
        U, S, VH, lost = res
    else:
        # how many singular triplets the truncation below can keep at most: chi_max; none of the deflated directions if they
        # are below `svd_min` anyway -- then their vectors need no orthonormal completion.  Deflated = values below
        # max(tol, rounding level) |theta|; the rounding-level threshold of the kernel is 16 eps sqrt(max(m, n)) per block
        # (1.6e-13 for 2048 columns), bounded here by 1e-11: a smaller `svd_min` may keep such directions (LAPACK reports
        # them as ~1e-17 |theta|, the reference's truncation keeps them down to svd_min) and they are completed
        n_keep = chi_max
        svd_min, chi_min = trunc_par.get('svd_min', 1.e-14), trunc_par.get('chi_min', None)
        if svd_min and svd_min >= max(tol, 1.e-11) and not (chi_min and chi_min > 1):
            n_keep = 0
        U, S, VH = npc.svd(theta, full_matrices=False, compute_uv=True, qtotal_LR=qtotal_LR,
                           inner_labels=inner_labels, guess=guess, deflation_tol=tol,
                           n_keep=(n_keep if full_out is None else None))
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
