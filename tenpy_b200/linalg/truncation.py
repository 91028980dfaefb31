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


def svd_theta(theta, trunc_par, qtotal_LR=[None, None], inner_labels=['vR', 'vL'], guess=None, full_out=None):
    """SVD of the matrix `theta` and truncation (reference truncation.py:258).

    Returns ``(U, S, VH, err, renormalization)`` with ``theta ~= U diag(S * renormalization) VH``.
    Extensions: `guess` is handed to :func:`npc.svd` (warm start); if `full_out` is a list, the untruncated
    ``(U, VH)`` are appended to it (shallow copies, to be used as the next guess).  Option
    ``trunc_par['svd_deflation_tol']`` (default 1e-10): singular directions below that fraction of ``|theta|``
    are not iterated to convergence inside the Jacobi SVD -- their values are reported approximately
    (absolute error below the tolerance) and their vectors are an orthonormal completion; the state changes by
    at most that relative amount, the energy to second order in it.  Without `full_out` only the vectors a
    truncation can keep (``chi_max``) are completed; with it the bases are completed fully, because the next
    warm start needs a complete orthonormal basis."""
    U, S, VH = npc.svd(theta, full_matrices=False, compute_uv=True, qtotal_LR=qtotal_LR, inner_labels=inner_labels,
                       guess=guess, deflation_tol=trunc_par.get('svd_deflation_tol', 1.e-10),
                       n_keep=(trunc_par.get('chi_max', 100) if full_out is None else None))
    if full_out is not None:
        full_out.append((U.copy(deep=False), VH.copy(deep=False)))
    renormalization = np.linalg.norm(S)
    S = S / renormalization
    piv, new_norm, err = truncate(S, trunc_par)
    new_len_S = np.sum(piv, dtype=np.int_)
    if new_len_S * 100 < len(S) and (trunc_par.get('chi_max', 100) is None or
                                     new_len_S != trunc_par.get('chi_max', 100)):
        warnings.warn('Catastrophic reduction in chi: {0:d} -> {1:d}'.format(len(S), int(new_len_S)), stacklevel=2)
    S = S[piv] / new_norm
    renormalization *= new_norm
    U.iproject(piv, axes=1)
    VH.iproject(piv, axes=0)
    return U, S, VH, err, renormalization
