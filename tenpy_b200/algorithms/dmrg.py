"""Two-site DMRG on the B200-native tensor engine.

Host-side mirror of the reference driver ``tenpy/algorithms/dmrg.py`` (`run` :63, `DMRGEngine` :112,
`TwoSiteDMRGEngine` :846) and of the sweep logic it inherits from ``mps_common.Sweep`` (:60; `sweep` :345,
`get_sweep_schedule` :419, `prepare_update_local` :498, `update_env` :569) and `IterativeSweeps.run`
(:796).  Same options (``trunc_params``, ``lanczos_params``, ``combine``, ``mixer``, ``mixer_params``,
``chi_list``, ``max_E_err``, ``max_S_err``, ``min_sweeps``, ``max_sweeps``, ``N_sweeps_check``), same update
order, same statistics keys.  Each bond update is: effective-H matvecs inside Lanczos (grouped FP64
tensor-core GEMMs), batched block SVD, environment update -- all on the device; the host only steers.
"""
# Copyright (C) 2026 tenpy_b200 authors. Apache-2.0.

import logging
import time

import numpy as np

from ..linalg import np_conserved as npc
from ..linalg.krylov_based import LanczosGroundState
from ..linalg.truncation import svd_theta, TruncationError
from ..networks.mpo import MPOEnvironment
from .mps_common import TwoSiteH, DensityMatrixMixer

logger = logging.getLogger(__name__)

__all__ = ['run', 'TwoSiteDMRGEngine', 'chi_list', 'entropy']


def entropy(p, n=1):
    """von-Neumann / Renyi entropy of a probability vector (reference tools/math.py:66)."""
    p = np.asarray(p)
    p = p[p > 1.e-30]
    if n == 1:
        return float(-np.inner(np.log(p), p))
    if n == np.inf:
        return float(-np.log(np.max(p)))
    return float(np.log(np.sum(p**n)) / (1. - n))


def chi_list(chi_max, dchi=20, nsweeps=20):
    """Ramp of bond dimensions ``{sweep: chi}`` (reference dmrg.py:1142)."""
    warmup = int(chi_max / dchi)
    if warmup == 0:
        return {0: chi_max}
    res = {}
    for i in range(warmup):
        res[i * nsweeps] = (i + 1) * dchi
    if chi_max > warmup * dchi:
        res[warmup * nsweeps] = chi_max
    return res


def run(psi, model, options):
    """Run two-site DMRG; `psi` is optimised in place (reference dmrg.py:63).

    Returns a dict with ``E``, ``shelve``, ``bond_statistics``, ``sweep_statistics``."""
    engine = TwoSiteDMRGEngine(psi, model, options)
    E, _ = engine.run()
    return {'E': E, 'shelve': False, 'bond_statistics': engine.update_stats,
            'sweep_statistics': engine.sweep_stats}


class TwoSiteDMRGEngine:
    """Engine of the two-site DMRG (reference dmrg.py:846 on top of :112 and mps_common.py:60)."""
    EffectiveH = TwoSiteH
    DefaultMixer = DensityMatrixMixer
    n_optimize = 2

    def __init__(self, psi, model, options):
        self.psi = psi
        self.model = model
        self.options = options = dict(options or {})
        self.finite = True
        self.combine = options.get('combine', True)
        self.trunc_params = dict(options.get('trunc_params', {}))
        self.lanczos_params = dict(options.get('lanczos_params', {}))
        self.chi_list = options.get('chi_list', None)
        self.diag_method = options.get('diag_method', 'lanczos')
        if self.diag_method not in ('default', 'lanczos'):
            raise NotImplementedError('diag_method ' + repr(self.diag_method))
        self.N_sweeps_check = options.get('N_sweeps_check', 1)
        self.min_sweeps = options.get('min_sweeps', int(1.5 * self.N_sweeps_check))
        self.max_sweeps = options.get('max_sweeps', 1000)
        self.max_E_err = options.get('max_E_err', 1.e-8)
        self.max_S_err = options.get('max_S_err', 1.e-5)
        self.max_seconds = 3600 * options.get('max_hours', 24 * 365)
        self.E_tol_to_trunc = options.get('E_tol_to_trunc', None)
        self.sweeps = options.get('sweep_0', 0)
        self.time0 = time.time()
        self.mixer = None
        # warm start of the Jacobi SVD from the previous update of the same bond (extension):
        #   'subspace' (default): decompose theta inside the span of the previously kept isometry when the part
        #                outside is below the truncation tolerance (truncation.svd_theta), no extra memory;
        #   'full'     : rotate theta with the complete previous singular vector bases (2 (chi d)^2 doubles per bond);
        #   False      : cold start every time.
        ws = options.get('svd_warm_start', 'subspace')
        self.svd_warm_start = 'full' if ws is True else ws
        self._svd_guess = {}
        self.env = MPOEnvironment(psi, model.H_MPO, psi)
        self.eff_H = None
        self.i0 = 0
        self.move_right = True
        self.update_LP_RP = (True, False)
        self.E_trunc_list = []
        self.trunc_err_list = []
        self._entropy_approx = [None] * psi.L
        self.reset_stats()
        self.mixer_activate()

    # ------------------------------------------------------------------ statistics
    def reset_stats(self):
        self.update_stats = {'i0': [], 'age': [], 'E_total': [], 'N_lanczos': [], 'time': [], 'err': [],
                             'E_trunc': [], 'ov_change': []}
        self.sweep_stats = {'sweep': [], 'N_updates': [], 'E': [], 'S': [], 'time': [], 'max_trunc_err': [],
                            'max_E_trunc': [], 'max_chi': [], 'norm_err': []}
        self.shelve = False
        self.time0 = time.time()

    # ------------------------------------------------------------------ mixer
    def mixer_activate(self):
        """Reference mps_common.py:653."""
        Mixer_class = self.options.get('mixer', False)
        if not Mixer_class:
            return
        if Mixer_class is True or Mixer_class == 'DensityMatrixMixer':
            Mixer_class = self.DefaultMixer
        self.mixer = Mixer_class(self.options.get('mixer_params', {}), self.sweeps)

    def mixer_deactivate(self):
        self.mixer = None

    # ------------------------------------------------------------------ the run loop (IterativeSweeps.run, :796)
    def run(self):
        self.is_first = True
        while True:
            if self.stopping_criterion():
                break
            self.run_iteration()
        self.post_run_cleanup()
        return self.sweep_stats['E'][-1] if len(self.sweep_stats['E']) else None, self.psi

    def stopping_criterion(self):
        """Reference dmrg.py:376 `is_converged` + mps_common.py:869."""
        if self.sweeps >= self.max_sweeps:
            return True
        if time.time() - self.time0 > self.max_seconds:
            self.shelve = True
            return True
        if self.sweeps < self.min_sweeps or len(self.sweep_stats['E']) < 2:
            return False
        if self.mixer is not None:
            return False
        if any(isinstance(s, npc.Array) for s in self.psi._S):
            # a sweep with the mixer leaves 2-D bond matrices; one mixer-free sweep restores the diagonal form
            # (the reference does this with Sweep.mixer_cleanup, mps_common.py:693)
            return False
        E = self.sweep_stats['E']
        S = self.sweep_stats['S']
        Delta_E = (E[-1] - E[-2]) / self.N_sweeps_check
        Delta_S = (S[-1] - S[-2]) / self.N_sweeps_check
        return abs(Delta_E / max(abs(E[-1]), 1.)) < self.max_E_err and abs(Delta_S) < self.max_S_err

    def run_iteration(self):
        """Reference dmrg.py:219."""
        max_trunc_err = 0.
        max_E_trunc = 0.
        for _ in range(self.N_sweeps_check):
            max_trunc_err = max(max_trunc_err, self.sweep())
            if len(self.E_trunc_list):
                max_E_trunc = max(max_E_trunc, np.max(np.abs(self.E_trunc_list)))
        E = self.update_stats['E_total'][-1]
        S_all = [s for s in self._entropy_approx if s is not None]
        S = max(S_all) if S_all else 0.
        self.sweep_stats['sweep'].append(self.sweeps)
        self.sweep_stats['N_updates'].append(len(self.update_stats['i0']))
        self.sweep_stats['E'].append(E)
        self.sweep_stats['S'].append(S)
        self.sweep_stats['time'].append(time.time() - self.time0)
        self.sweep_stats['max_trunc_err'].append(max_trunc_err)
        self.sweep_stats['max_E_trunc'].append(max_E_trunc)
        self.sweep_stats['max_chi'].append(int(np.max(self.psi.chi)))
        logger.info('sweep %d: E=%.13f S=%.6f chi=%d trunc=%.2e t=%.1fs', self.sweeps, E, S,
                    self.sweep_stats['max_chi'][-1], max_trunc_err, self.sweep_stats['time'][-1])
        return E, self.psi

    def post_run_cleanup(self):
        pass

    # ------------------------------------------------------------------ one sweep (mps_common.py:345)
    def get_sweep_schedule(self):
        L, n = self.psi.L, 2
        assert L > n
        i0s = list(range(0, L - n)) + list(range(L - n, 0, -1))
        move_right = [True] * (L - n) + [False] * (L - n)
        update_LP_RP = [[True, False]] * (L - n) + [[False, True]] * (L - n)
        return zip(i0s, move_right, update_LP_RP)

    def sweep(self, optimize=True):
        self.E_trunc_list = []
        self.trunc_err_list = []
        if optimize and self.chi_list is not None:
            new_chi_max = self.chi_list.get(self.sweeps, None)
            if new_chi_max is not None:
                self.trunc_params['chi_max'] = new_chi_max
                if self.options.get('chi_list_reactivates_mixer', True):
                    self.mixer_activate()
        for i0, move_right, update_LP_RP in self.get_sweep_schedule():
            self.i0, self.move_right, self.update_LP_RP = i0, move_right, update_LP_RP
            theta = self.prepare_update_local()
            update_data = self.update_local(theta, optimize=optimize)
            self.update_env(**update_data)
            self.post_update_local(**update_data)
            self.free_no_longer_needed_envs()
        if optimize:
            self.sweeps += 1
            if self.mixer is not None:
                mixer = self.mixer.update_amplitude(self.sweeps)
                if mixer is None:
                    self.mixer_deactivate()
                else:
                    self.mixer = mixer
        return np.max(self.trunc_err_list)

    def prepare_update_local(self):
        """Reference mps_common.py:498."""
        self.eff_H = self.EffectiveH(self.env, self.i0, self.combine, self.move_right)
        theta = self.psi.get_theta(self.i0, n=2)
        return self.eff_H.combine_theta(theta)

    def update_local(self, theta, optimize=True):
        """Reference dmrg.py:529."""
        i0 = self.i0
        age = self.env.get_LP_age(i0) + 2 + self.env.get_RP_age(i0 + 1)
        if optimize:
            E0, theta, N, ov_change = self.diag(theta)
        else:
            E0, N, ov_change = None, 0, 0.
        theta = self.prepare_svd(theta)
        U, S, VH, err, S_approx = self.mixed_svd(theta)
        self._entropy_approx[i0 + 1] = entropy(np.asarray(S_approx)**2)
        self.set_B(U, S, VH)
        return {'E0': E0, 'err': err, 'N': N, 'age': age, 'U': U, 'VH': VH, 'ov_change': ov_change}

    def diag(self, theta_guess):
        """Reference dmrg.py:672 (Lanczos only; the reference's small-N ED shortcut is a host LAPACK call)."""
        E, theta, N = LanczosGroundState(self.eff_H, theta_guess, self.lanczos_params).run()
        ov_change = 1. - abs(npc.inner(theta_guess, theta, 'labels', do_conj=True))
        return E, theta, N, ov_change

    def prepare_svd(self, theta):
        if self.combine:
            return theta
        return theta.combine_legs([['vL', 'p0'], ['p1', 'vR']], new_axes=[0, 1], qconj=[+1, -1])

    def mixed_svd(self, theta):
        """Reference dmrg.py:876."""
        i0 = self.i0
        update_LP, update_RP = self.update_LP_RP
        if self.mixer is None:
            qtotal_i0 = self.psi.get_B(i0, form=None).qtotal
            ws = self.svd_warm_start
            full = [] if ws == 'full' else None
            U, S, VH, err, _ = svd_theta(theta, self.trunc_params, qtotal_LR=[qtotal_i0, None],
                                         inner_labels=['vR', 'vL'],
                                         guess=self._svd_guess.get(i0) if ws == 'full' else None, full_out=full,
                                         subspace=self._svd_guess.get(i0) if ws == 'subspace' else None)
            if full:
                self._svd_guess[i0] = full[0]
            elif ws == 'subspace':
                self._svd_guess[i0] = (U.copy(deep=False), VH.copy(deep=False))
            S_a = S
        else:
            old_BL_qtotal = self.psi.get_B(i0, form=None).qtotal
            qtotal_LR = [old_BL_qtotal, theta.chinfo.make_valid(theta.qtotal - old_BL_qtotal)]
            U, S, VH, err, S_a = self.mixer.mix_and_decompose_2site(engine=self, theta=theta, i0=i0,
                                                                    mix_left=update_LP, mix_right=update_RP,
                                                                    qtotal_LR=qtotal_LR)
        U.ireplace_label('(vL.p0)', '(vL.p)')
        VH.ireplace_label('(p1.vR)', '(p.vR)')
        return U, S, VH, err, S_a

    def set_B(self, U, S, VH):
        """Reference dmrg.py:934."""
        B0 = U.split_legs(['(vL.p)'])
        B1 = VH.split_legs(['(p.vR)'])
        i0 = self.i0
        self.psi.set_B(i0, B0, form='A')
        self.psi.set_B(i0 + 1, B1, form='B')
        self.psi.set_SR(i0, S)

    def update_env(self, **update_data):
        """Reference mps_common.py:569."""
        i0 = self.i0
        update_LP, update_RP = self.update_LP_RP
        if update_LP:
            self.eff_H.update_LP(self.env, i0 + 1, update_data['U'])
        if update_RP:
            self.eff_H.update_RP(self.env, i0, update_data['VH'])

    def post_update_local(self, E0, age, N, ov_change, err, **update_data):
        """Reference dmrg.py:575."""
        i0 = self.i0
        E_trunc = None
        self.trunc_err_list.append(err.eps)
        self.update_stats['i0'].append(i0)
        self.update_stats['age'].append(age)
        self.update_stats['E_total'].append(E0)
        self.update_stats['E_trunc'].append(E_trunc)
        self.update_stats['N_lanczos'].append(N)
        self.update_stats['err'].append(err)
        self.update_stats['ov_change'].append(ov_change)
        self.update_stats['time'].append(time.time() - self.time0)

    def free_no_longer_needed_envs(self):
        """Reference mps_common.py:614: parts that will be recomputed before their next use are dropped."""
        i0 = self.i0
        update_LP, update_RP = self.update_LP_RP
        if update_LP and not update_RP:
            # moving right: RP[i0] is outdated (site i0+1 changed)
            if i0 < self.psi.L - 1:
                self.env.del_RP(i0)
        if update_RP and not update_LP:
            if i0 + 1 > 0:
                self.env.del_LP(i0 + 1)
