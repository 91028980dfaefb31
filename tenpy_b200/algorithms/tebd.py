"""TEBD bond updates on the B200-native tensor engine (the next caller of the hot path).

Host-side mirror of the reference's ``tenpy/algorithms/tebd.py``: `TEBDEngine.calc_U` (:297) / `_calc_U_bond`
(:585), `suzuki_trotter_time_steps` (:183), `suzuki_trotter_decomposition` (:219), `evolve` (:346),
`evolve_step` (:374), `update_bond` (:416), `update_imag` (:485), `update_bond_imag` (:545) and `run_GS` (:113).

A bond update needs **no new kernel**: ``tensordot(U_bond, theta)`` is a small-k grouped GEMM, the leg fusion a
strided block copy, the truncation the batched block-Jacobi SVD behind `svd_theta`, and the ``C V^dagger`` trick
of the reference (no inverse Schmidt values) another grouped GEMM.  The engine computes in float64, so only
``type_evo='imag'`` is available (real-time evolution needs complex Arrays: out of scope of round 1).
"""
# Copyright (C) 2026 tenpy_b200 authors. Apache-2.0.

import numpy as np
import scipy.linalg

from ..linalg import np_conserved as npc
from ..linalg.truncation import svd_theta, TruncationError, decompose_theta_qr_based

__all__ = ['TEBDEngine', 'QRBasedTEBDEngine']


class TEBDEngine:
    """Time evolving block decimation (reference tebd.py:43 `TEBDEngine`).

    Options (same names as the reference): ``trunc_params``, ``order``, ``N_steps``, ``delta_tau_list``,
    ``max_error_E``.  `model` provides ``H_bond`` and ``bond_energies`` (see :mod:`tenpy_b200.models`)."""

    def __init__(self, psi, model, options):
        self.psi = psi
        self.model = model
        self.options = options = dict(options or {})
        self.trunc_params = dict(options.get('trunc_params', {}))
        self.trunc_err = TruncationError()
        self._trunc_err_bonds = [TruncationError() for _ in range(psi.L + 1)]
        self._U = None
        self._U_param = {}
        self._update_index = None
        self.evolved_time = 0.

    @property
    def trunc_err_bonds(self):
        """truncation error introduced on each non-trivial bond"""
        return self._trunc_err_bonds[1:self.psi.L]

    # ------------------------------------------------------------------ Trotter decomposition
    @staticmethod
    def suzuki_trotter_time_steps(order):
        """fractions of ``delta_t`` for which ``exp(-H_bond dt)`` is needed (reference tebd.py:183)"""
        if order == 1:
            return [1.]
        if order == 2:
            return [0.5, 1.]
        if order == 4:
            t1 = 1. / (4. - 4.**(1. / 3.))
            t3 = 1. - 4. * t1
            return [t1 / 2., t1, (t1 + t3) / 2., t3]
        raise ValueError('Unknown order {0!r} for Suzuki Trotter decomposition'.format(order))

    @staticmethod
    def suzuki_trotter_decomposition(order, N_steps):
        """list of ``(time step index, odd/even)`` pairs for `N_steps` steps (reference tebd.py:219)"""
        even, odd = 0, 1
        if N_steps == 0:
            return []
        if order == 1:
            return [(0, odd), (0, even)] * N_steps
        if order == 2:
            a, a2, b = (0, odd), (1, odd), (1, even)
            return [a, b] + [a2, b] * (N_steps - 1) + [a]
        if order == 4:
            a, a2, b = (0, odd), (1, odd), (1, even)
            c, d = (2, odd), (3, even)
            # one step is U(t1) U(t1) U(t3) U(t1) U(t1) with U(t) = odd(t/2) even(t) odd(t/2); neighbouring odd
            # half layers are fused, also across consecutive steps (the trailing `a` of one step and the leading
            # `a` of the next one become `a2`)
            body = [b, a2, b, c, d, c, b, a2, b]
            steps = [a] + body
            for _ in range(N_steps - 1):
                steps += [a2] + body
            return steps + [a]
        raise ValueError('Unknown order {0!r} for Suzuki Trotter decomposition'.format(order))

    def calc_U(self, order, delta_t, type_evo='imag', E_offset=None):
        """``self._U[k][i] = exp(-dt_k H_bond[i])`` for every Trotter sub-step (reference tebd.py:297)."""
        if type_evo != 'imag':
            raise NotImplementedError("type_evo='real' needs complex Arrays; this engine computes in float64")
        U_param = dict(order=order, delta_t=delta_t, type_evo=type_evo, E_offset=E_offset, tau=delta_t)
        if self._U_param == U_param:
            return
        self._U_param = U_param
        L = self.psi.L
        self._U = []
        for dt in self.suzuki_trotter_time_steps(order):
            self._U.append([self._calc_U_bond(i_bond, dt * delta_t, type_evo, E_offset) for i_bond in range(L)])

    def _calc_U_bond(self, i_bond, dt, type_evo, E_offset):
        """``exp(-dt H_bond)`` (reference tebd.py:585).  The ``d^2 x d^2`` exponential is host work, like the
        reference's `npc.expm` (a scipy call per charge block); the result is charge conserving because
        `H_bond` is."""
        h = self.model.H_bond[i_bond]
        if h is None:
            return None
        hd = h.to_ndarray()                                   # p0, p0*, p1, p1*
        d0, d1 = hd.shape[0], hd.shape[2]
        H2 = hd.transpose(0, 2, 1, 3).reshape(d0 * d1, d0 * d1)
        U = scipy.linalg.expm(-dt * H2).reshape(d0, d1, d0, d1)
        legs = [h.get_leg('p0'), h.get_leg('p1'), h.get_leg('p0*'), h.get_leg('p1*')]
        return npc.Array.from_ndarray(U, legs, labels=['p0', 'p1', 'p0*', 'p1*'], cutoff=1e-16)

    # ------------------------------------------------------------------ brick-wall evolution
    def evolve(self, N_steps, dt):
        """`N_steps` Trotter steps with the prepared `U` (reference tebd.py:346)."""
        if dt is not None:
            assert dt == self._U_param['delta_t']
        trunc_err = TruncationError()
        for U_idx_dt, odd in self.suzuki_trotter_decomposition(self._U_param['order'], N_steps):
            trunc_err += self.evolve_step(U_idx_dt, odd)
        self.evolved_time = self.evolved_time + N_steps * self._U_param['tau']
        self.trunc_err = self.trunc_err + trunc_err
        return trunc_err

    def evolve_step(self, U_idx_dt, odd):
        """update all even or all odd bonds (reference tebd.py:374)"""
        Us = self._U[U_idx_dt]
        trunc_err = TruncationError()
        for i_bond in np.arange(int(odd) % 2, self.psi.L, 2):
            if Us[i_bond] is None:
                continue
            self._update_index = (U_idx_dt, i_bond)
            trunc_err += self.update_bond(i_bond, Us[i_bond])
        self._update_index = None
        return trunc_err

    def update_bond(self, i, U_bond):
        """Apply `U_bond` to sites ``(i-1, i)``, truncate, keep both in ``'B'`` form (reference tebd.py:416-483)."""
        i0, i1 = i - 1, i
        C = self.psi.get_theta(i0, n=2, formL=0.)             # the two B without the S on the left
        C = npc.tensordot(U_bond, C, axes=(['p0*', 'p1*'], ['p0', 'p1']))
        C.itranspose(['vL', 'p0', 'p1', 'vR'])
        theta = C.scale_axis(self.psi.get_SL(i0), 'vL')
        theta = theta.combine_legs([('vL', 'p0'), ('p1', 'vR')], qconj=[+1, -1])
        U, S, V, trunc_err, renormalize = svd_theta(theta, self.trunc_params,
                                                    [self.psi.get_B(i0, None).qtotal, None],
                                                    inner_labels=['vR', 'vL'])
        B_R = V.split_legs(1).ireplace_label('p1', 'p')
        # B_L = SL^-1 U S = SL^-1 U S V V^dagger = C V^dagger: no inverse of small Schmidt values (tebd.py:464-476)
        B_L = npc.tensordot(C.combine_legs(('p1', 'vR'), pipes=theta.legs[1]), V.conj(),
                            axes=['(p1.vR)', '(p1*.vR*)'])
        B_L.ireplace_labels(['vL*', 'p0'], ['vR', 'p'])
        B_L /= renormalize
        self.psi.norm *= renormalize
        self.psi.set_SR(i0, S)
        self.psi.set_B(i0, B_L, form='B')
        self.psi.set_B(i1, B_R, form='B')
        self._trunc_err_bonds[i] = self._trunc_err_bonds[i] + trunc_err
        return trunc_err

    # ------------------------------------------------------------------ imaginary time: sweeps
    def update_imag(self, N_steps, call_canonical_form=False):
        """Imaginary-time update by sweeping right and left with half steps, like DMRG, which preserves the
        orthonormality of the canonical form (reference tebd.py:485-543).  Second order, finite chains only."""
        if call_canonical_form:
            raise NotImplementedError('MPS.canonical_form is outside the hot path (SURVEY.md section 8)')
        trunc_err = TruncationError()
        order = self._U_param['order']
        if order != 2 or not self.psi.finite:
            raise NotImplementedError('Use DMRG instead...')
        U_idx_dt = 0
        assert self.suzuki_trotter_time_steps(order)[U_idx_dt] == 0.5
        Us = self._U[U_idx_dt]
        for _ in range(N_steps):
            for i_bond in range(self.psi.L):                  # sweep right
                if Us[i_bond] is None:
                    continue
                self._update_index = (U_idx_dt, i_bond)
                trunc_err += self.update_bond_imag(i_bond, Us[i_bond])
            for i_bond in range(self.psi.L - 1, -1, -1):      # sweep left
                if Us[i_bond] is None:
                    continue
                self._update_index = (U_idx_dt, i_bond)
                trunc_err += self.update_bond_imag(i_bond, Us[i_bond])
        self._update_index = None
        self.evolved_time = self.evolved_time + N_steps * self._U_param['tau']
        self.trunc_err = self.trunc_err + trunc_err
        return trunc_err

    def update_bond_imag(self, i, U_bond):
        """Update with a non-unitary `U_bond`, keeping ``A S B`` around the bond (reference tebd.py:545-583)."""
        i0, i1 = i - 1, i
        theta = self.psi.get_theta(i0, n=2)
        theta = npc.tensordot(U_bond, theta, axes=(['p0*', 'p1*'], ['p0', 'p1']))
        theta = theta.combine_legs([('vL', 'p0'), ('vR', 'p1')], qconj=[+1, -1])
        U, S, V, trunc_err, renormalize = svd_theta(theta, self.trunc_params, inner_labels=['vR', 'vL'])
        self.psi.norm *= renormalize
        B_R = V.split_legs(1).ireplace_label('p1', 'p')
        A_L = U.split_legs(0).ireplace_label('p0', 'p')
        self.psi.set_SR(i0, S)
        self.psi.set_B(i0, A_L, form='A')
        self.psi.set_B(i1, B_R, form='B')
        self._trunc_err_bonds[i] = self._trunc_err_bonds[i] + trunc_err
        return trunc_err

    # ------------------------------------------------------------------ ground state search
    def run_GS(self):
        """Imaginary time evolution with decreasing time steps until the bond energy is converged (reference
        tebd.py:113-180).  Returns the last mean bond energy."""
        delta_tau_list = self.options.get('delta_tau_list', [0.1, 0.01, 0.001, 1.e-4, 1.e-5, 1.e-6, 1.e-7, 1.e-8,
                                                             1.e-9, 1.e-10, 1.e-11, 0.])
        max_error_E = self.options.get('max_error_E', 1.e-13)
        N_steps = self.options.get('N_steps', 10)
        order = self.options.get('order', 2)
        Eold = np.mean(self.model.bond_energies(self.psi))
        for delta_tau in delta_tau_list:
            self.calc_U(order, delta_tau, type_evo='imag')
            DeltaE = 2 * max_error_E
            while DeltaE > max_error_E:
                if self.psi.finite and order == 2:
                    self.update_imag(N_steps, call_canonical_form=False)
                else:
                    self.evolve(N_steps, delta_tau)
                E = np.mean(self.model.bond_energies(self.psi))
                DeltaE = abs(Eold - E)
                Eold = E
        return Eold


class QRBasedTEBDEngine(TEBDEngine):
    """TEBD with the QR based truncation of :func:`~tenpy_b200.linalg.truncation.decompose_theta_qr_based` instead of the
    SVD of the full two-site wave function (reference tebd.py:619, arXiv:2212.09782): only the small bond matrix is
    decomposed.  Options as the reference: `cbe_expand` (0.1), `cbe_expand_0`, `cbe_min_block_increase` (1),
    `use_eig_based_svd` (False), `compute_err` (True)."""

    def _expansion_rate(self, i):
        """Reference tebd.py:669."""
        expand = self.options.get('cbe_expand', 0.1)
        expand_0 = self.options.get('cbe_expand_0', None)
        if expand_0 is None or expand_0 == expand:
            return expand
        chi_max = self.trunc_params.get('chi_max', None)
        if chi_max is None:
            raise ValueError('Need to specify trunc_params["chi_max"] in order to use cbe_expand_0.')
        chi = min(np.shape(self.psi.get_SL(i)))
        return max(expand_0 - chi / chi_max * (expand_0 - expand), expand)

    def _two_site_theta(self, i0, U_bond, with_left_S):
        """`U_bond` applied to sites ``(i0, i0+1)``; returns ``(C, theta)``: `C` without, `theta` (combined to a
        matrix) with the Schmidt values on the left bond (`C` is None for ``with_left_S=True``)."""
        C = self.psi.get_theta(i0, n=2, formL=1. if with_left_S else 0.)
        C = npc.tensordot(U_bond, C, axes=(['p0*', 'p1*'], ['p0', 'p1']))
        C.itranspose(['vL', 'p0', 'p1', 'vR'])
        theta = C if with_left_S else C.scale_axis(self.psi.get_SL(i0), 'vL')
        theta = theta.combine_legs([('vL', 'p0'), ('p1', 'vR')], qconj=[+1, -1])
        return (None if with_left_S else C), theta

    def _qr_split(self, i0, theta, both, eig_ok=True):
        """:func:`decompose_theta_qr_based` of `theta` on bond ``(i0, i0+1)`` with this engine's options."""
        old_L, old_R = self.psi.get_B(i0, 'B'), self.psi.get_B(i0 + 1, 'B')
        use_eig = self.options.get('use_eig_based_svd', False)
        if use_eig and not eig_ok:
            raise NotImplementedError('update_bond_imag does not (yet) support eig based SVD')
        return decompose_theta_qr_based(
            old_qtotal_L=old_L.qtotal, old_qtotal_R=old_R.qtotal, old_bond_leg=old_R.get_leg('vL'), theta=theta,
            move_right=False, expand=self._expansion_rate(i0 + 1),
            min_block_increase=self.options.get('cbe_min_block_increase', 1), use_eig_based_svd=use_eig,
            trunc_params=self.trunc_params, compute_err=self.options.get('compute_err', True), return_both_T=both)

    def _store(self, i0, left, left_form, S, B_R, renormalize, trunc_err):
        self.psi.norm *= renormalize
        self.psi.set_B(i0, left, form=left_form)
        self.psi.set_SL(i0 + 1, S)
        self.psi.set_B(i0 + 1, B_R.split_legs(1), form='B')
        self._trunc_err_bonds[i0 + 1] = self._trunc_err_bonds[i0 + 1] + trunc_err
        return trunc_err

    def update_bond(self, i, U_bond):
        """Unitary-style update keeping both sites in 'B' form (reference tebd.py:681): only the right isometry comes
        out of the QR based split, ``B_L = C . B_R^dagger`` avoids dividing by small Schmidt values."""
        i0 = i - 1
        C, theta = self._two_site_theta(i0, U_bond, with_left_S=False)
        _, S, B_R, form, trunc_err, renormalize = self._qr_split(i0, theta, both=False)
        assert form[1] == 'B'
        B_L = npc.tensordot(C.combine_legs(('p1', 'vR'), pipes=theta.legs[1]), B_R.conj(),
                            axes=['(p1.vR)', '(p*.vR*)']) / renormalize
        B_L.ireplace_labels(['p0', 'vL*'], ['p', 'vR'])
        return self._store(i0, B_L, 'B', S, B_R, renormalize, trunc_err)

    def update_bond_imag(self, i, U_bond):
        """Non-unitary update leaving ``A S B`` around the bond (reference tebd.py:741)."""
        i0 = i - 1
        _, theta = self._two_site_theta(i0, U_bond, with_left_S=True)
        A_L, S, B_R, form, trunc_err, renormalize = self._qr_split(i0, theta, both=True, eig_ok=False)
        assert form == ['A', 'B']
        return self._store(i0, A_L.split_legs(0), 'A', S, B_R, renormalize, trunc_err)
