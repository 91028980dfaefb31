"""ORACLE (test infrastructure, CPU only) -- dense numpy/LAPACK restatement of the two-site DMRG hot path.

This file is NOT part of the product: only ``tests/``, ``__graft_entry__.smoke()`` and the CPU-baseline /
``--impl reference`` legs of ``bench.py`` may import it.  It restates, for tensors WITHOUT charge
conservation (every Array is one dense block; BASELINE.json configs 1 and 2), what the reference computes
with NumPy/OpenBLAS/LAPACK, function by function:

* `matvec`            <- TwoSiteH.matvec, combine=True branch   (tenpy/algorithms/mps_common.py:1337-1339)
* `matvec_split`      <- TwoSiteH.matvec, combine=False branch  (mps_common.py:1340-1350)
* `lanczos_ground`    <- LanczosGroundState.run / _build_krylov / _converged / _calc_result_full
                         (tenpy/linalg/krylov_based.py:614, :645, :677, :160)
* `truncate`          <- truncation.truncate                     (tenpy/linalg/truncation.py:146)
* `svd_theta`         <- truncation.svd_theta -> npc.svd -> svd_robust.svd (LAPACK gesdd, fallback gesvd)
                         (truncation.py:258, np_conserved.py:4950, svd_robust.py:37)
* `contract_LHeff/RHeff`, `update_LP/RP` <- MPOEnvironment._contract_LHeff/_RHeff (networks/mpo.py:3107,
                         :3118), TwoSiteH.update_LP/update_RP (mps_common.py:1421, :1430)
* `run_dmrg`          <- Sweep.sweep schedule (mps_common.py:345-419) + DMRGEngine.update_local
                         (dmrg.py:529) with mixer=None, combine=True, diag_method='lanczos'.

Pinned: `run_dmrg` reproduces the reference energy of BASELINE.md config 1
(TFIChain L=20, chi_max=50: E = -25.1077971116238) and the golden vectors in tests/golden (generated from
the reference by tests/golden/make_golden.py); see tests/test_oracle.py.
"""
import numpy as np
import scipy.linalg


# ----------------------------------------------------------------------------- truncation (truncation.py:146)
def truncate(S, chi_max=100, chi_min=None, degeneracy_tol=None, svd_min=1e-14, trunc_cut=1e-14):
    S = np.asarray(S)
    logS = np.log(np.choose(S <= 0., [S, 1e-100 * np.ones(len(S))]))
    piv = np.argsort(logS)
    logS = logS[piv]
    good = np.ones(len(piv), dtype=bool)

    def comb(g1, g2):
        r = np.logical_and(g1, g2)
        return r if np.any(r) else g1
    if chi_max is not None:
        g2 = np.zeros(len(piv), dtype=bool)
        g2[-chi_max:] = True
        good = comb(good, g2)
    if chi_min is not None and chi_min > 1:
        g2 = np.ones(len(piv), dtype=bool)
        g2[-chi_min + 1:] = False
        good = comb(good, g2)
    if degeneracy_tol:
        g2 = np.empty(len(piv), bool)
        g2[0] = True
        g2[1:] = np.greater_equal(logS[1:] - logS[:-1], degeneracy_tol)
        good = comb(good, g2)
    if svd_min is not None:
        good = comb(good, np.greater_equal(logS, np.log(svd_min)))
    if trunc_cut is not None:
        good = comb(good, np.cumsum(S[piv]**2) > trunc_cut * trunc_cut)
    cut = np.nonzero(good)[0][0]
    mask = np.zeros(len(S), dtype=bool)
    mask[piv[cut:]] = True
    return mask, np.linalg.norm(S[mask]), np.sum(S[~mask]**2)


def svd_flat(a):
    """svd_robust.svd (svd_robust.py:37): gesdd, fallback gesvd"""
    try:
        return scipy.linalg.svd(a, full_matrices=False, lapack_driver='gesdd')
    except np.linalg.LinAlgError:
        return scipy.linalg.svd(a, full_matrices=False, lapack_driver='gesvd')


def svd_theta(theta, trunc_par):
    """truncation.svd_theta (truncation.py:258) for a dense matrix; returns U, S, VH, err, renormalization"""
    U, S, VH = svd_flat(theta)
    renorm = np.linalg.norm(S)
    S = S / renorm
    mask, new_norm, err = truncate(S, **trunc_par)
    S = S[mask] / new_norm
    return U[:, mask], S, VH[mask, :], err, renorm * new_norm


# ----------------------------------------------------------------------------- effective H (mps_common.py:1321)
def matvec(LHeff, RHeff, theta):
    """LHeff (n, D, n) [(vR*.p0), wR, (vR.p0*)], RHeff (D, n, n) [wL, (p1*.vL), (p1.vL*)], theta (n, n)."""
    t = np.tensordot(LHeff, theta, axes=[2, 0])            # (n, D, n)
    return np.tensordot(t, RHeff, axes=[[1, 2], [0, 1]])   # (n, n)


def matvec_split(LP, W0, W1, RP, theta4):
    """The reference's ``combine=False`` branch of TwoSiteH.matvec (mps_common.py:1340-1350): LP, W0, W1, RP one
    after the other.  LP (chi, D, chi) [vR*, wR, vR], W [wL, wR, p, p*], RP (chi, D, chi) [vL, wL, vL*],
    theta4 (chi, d, d, chi) [vL, p0, p1, vR]; returns [vL, p0, p1, vR].  Same result as `matvec` on the combined
    tensors with d times fewer flops in the chi^3 terms."""
    t = np.tensordot(LP, theta4, axes=[2, 0])               # vR*, wR, p0, p1, vR
    t = np.tensordot(W0, t, axes=[[0, 3], [1, 2]])          # wR, p0, vR*, p1, vR
    t = np.tensordot(t, W1, axes=[[0, 3], [0, 3]])          # p0, vR*, vR, wR, p1
    t = np.tensordot(t, RP, axes=[[3, 2], [1, 0]])          # p0, vR*, p1, vL*
    return t.transpose(1, 0, 2, 3)


def matvec_flops(n_left, D, n_right):
    """flops of one matvec = sum 2 m k n over the two GEMMs"""
    return 2. * (n_left * D) * n_left * n_right + 2. * n_left * (D * n_right) * n_right


# ----------------------------------------------------------------------------- Lanczos (krylov_based.py:584)
def lanczos_ground(matvec_fn, psi0, N_min=2, N_max=20, P_tol=1e-14, E_tol=np.inf, min_gap=1e-12, cutoff=None):
    if cutoff is None:
        cutoff = np.finfo(np.float64).eps * 100
    h = np.zeros((N_max + 1, N_max + 1))
    Es = np.zeros((N_max, N_max))
    cache = []
    w = psi0.copy()
    beta = np.linalg.norm(w)
    psi0n = None
    vk = np.ones(1)
    k = 0
    for k in range(N_max):
        w = w / beta
        if psi0n is None:
            psi0n = w
        cache.append(w)
        w = matvec_fn(w)
        alpha = float(np.sum(w * cache[-1]))
        h[k, k] = alpha
        if k == 0:
            Es[0, 0] = alpha
            vk = np.ones(1)
        else:
            E_kr, v_kr = np.linalg.eigh(h[:k + 1, :k + 1])
            Es[k, :k + 1] = E_kr
            vk = v_kr[:, 0]
        w = w - alpha * cache[-1]
        if k > 0:
            w = w - beta * cache[-2]
        beta = np.linalg.norm(w)
        h[k, k + 1] = h[k + 1, k] = beta
        if abs(beta) < cutoff:
            break
        if k + 1 >= N_min:
            RitzRes = abs(vk[k]) * h[k, k + 1]
            gap = max(Es[k, 1] - Es[k, 0], min_gap)
            if (RitzRes / gap)**2 < P_tol and Es[k - 1, 0] - Es[k, 0] < E_tol:
                break
    N = k + 1
    E0 = Es[N - 1, 0]
    if N == 1:
        return E0, psi0n.copy(), N
    psif = psi0n * vk[0]
    for j in range(1, N):
        psif = psif + vk[j] * cache[j]
    return E0, psif / np.linalg.norm(psif), N


# ----------------------------------------------------------------------------- environments (mpo.py:3087-3126)
def contract_LHeff(LP, W):
    """LP (chi, D, chi) [vR*, wR, vR], W (D, D, d, d) [wL, wR, p, p*] -> (chi d, D, chi d)"""
    t = np.tensordot(LP, W, axes=[1, 0])                   # vR*, vR, wR, p, p*
    t = t.transpose(0, 3, 2, 1, 4)                         # vR*, p, wR, vR, p*
    s = t.shape
    return t.reshape(s[0] * s[1], s[2], s[3] * s[4])


def contract_RHeff(RP, W):
    """RP (chi, D, chi) [vL, wL, vL*], W [wL, wR, p, p*] -> (D, d chi, d chi) [wL, (p1*.vL), (p1.vL*)]"""
    t = np.tensordot(W, RP, axes=[1, 1])                   # wL, p, p*, vL, vL*
    t = t.transpose(0, 2, 3, 1, 4)                         # wL, p*, vL, p, vL*
    s = t.shape
    return t.reshape(s[0], s[1] * s[2], s[3] * s[4])


def update_LP(LHeff, U):
    """LP' = U^dagger (LHeff U)  (mps_common.py:1421); U (n, chi')"""
    t = np.tensordot(LHeff, U, axes=[2, 0])                # (n, D, chi')
    return np.tensordot(U.conj(), t, axes=[0, 0])          # (chi', D, chi')


def update_RP(RHeff, VH):
    """RP' = (VH RHeff) VH^dagger  (mps_common.py:1430); VH (chi', n); result [vL, wL, vL*]"""
    t = np.tensordot(VH, RHeff, axes=[1, 1])               # (chi', D, n)
    return np.tensordot(t, VH.conj(), axes=[2, 1])         # (chi', D, chi')


def bond_update(LHeff, RHeff, theta, trunc_par, lanczos_par, move_right=True, matvec_fn=None):
    """One two-site update on dense tensors: Lanczos -> svd_theta -> environment update (dmrg.py:529).

    `matvec_fn` (optional) replaces the combined matvec inside Lanczos (e.g. the ``combine=False`` order).
    Returns (E0, U, S, VH, new environment part, N_lanczos)."""
    if matvec_fn is None:
        def matvec_fn(x):
            return matvec(LHeff, RHeff, x)
    E0, th, N = lanczos_ground(matvec_fn, theta, **lanczos_par)
    U, S, VH, err, _ = svd_theta(th, trunc_par)
    env = update_LP(LHeff, U) if move_right else update_RP(RHeff, VH)
    return E0, U, S, VH, env, N


# ----------------------------------------------------------------------------- whole finite DMRG (no charges)
def tfi_mpo(g=1., J=1.):
    """W[wL, wR, p, p*] of H = -J sum sx sx - g sum sz (reference tf_ising.py:74), D=3"""
    sx = np.array([[0., 1.], [1., 0.]])
    sz = np.array([[1., 0.], [0., -1.]])
    idm = np.eye(2)
    W = np.zeros((3, 3, 2, 2))
    W[0, 0] = idm
    W[0, 1] = sx
    W[0, 2] = -g * sz
    W[1, 2] = -J * sx
    W[2, 2] = idm
    return W


def run_dmrg(W, L, d, p_state, trunc_par, lanczos_par=None, max_sweeps=30, max_E_err=1e-10, max_S_err=1e-5,
             min_sweeps=1):
    """Finite two-site DMRG for a translation invariant MPO `W` (IdL=0, IdR=D-1), product initial state.

    Follows the reference driver with combine=True, mixer=None: schedule mps_common.py:419, update_local
    dmrg.py:529, convergence dmrg.py:376.  Returns dict(E, S (list per bond), Bs, sweeps, E_sweeps)."""
    lanczos_par = dict(lanczos_par or {})
    D = W.shape[0]
    Bs = []
    for i in range(L):
        B = np.zeros((1, d, 1))
        B[0, p_state[i], 0] = 1.
        Bs.append(B)
    Ss = [np.ones(1) for _ in range(L + 1)]
    form = ['B'] * L
    LP = [None] * L
    RP = [None] * L
    LP[0] = np.zeros((1, D, 1))
    LP[0][0, 0, 0] = 1.
    RP[L - 1] = np.zeros((1, D, 1))
    RP[L - 1][0, D - 1, 0] = 1.
    for i in range(L - 1, 1, -1):   # RP[i-1] from RP[i]  (mpo.py:3097)
        B = Bs[i]
        t = np.tensordot(B, RP[i], axes=[2, 0])                      # vL, p, wL, vL*
        t = np.tensordot(t, W, axes=[[1, 2], [3, 1]])                # vL, vL*, wL, p
        RP[i - 1] = np.tensordot(t, B.conj(), axes=[[3, 1], [1, 2]])  # vL, wL, vL*
    E_sweeps, S_sweeps = [], []
    sweeps = 0
    i0s = list(range(0, L - 2)) + list(range(L - 2, 0, -1))
    moves = [True] * (L - 2) + [False] * (L - 2)
    E0 = None
    Sbond = [0.] * (L + 1)
    while sweeps < max_sweeps:
        for i0, mr in zip(i0s, moves):
            LHeff = contract_LHeff(LP[i0], W)
            RHeff = contract_RHeff(RP[i0 + 1], W)
            # theta = S^1 B B or A S B etc: get_theta(i0, 2)  (mps.py:3041)
            B0, B1 = Bs[i0], Bs[i0 + 1]
            if form[i0] == 'B':
                B0 = Ss[i0][:, None, None] * B0
            if form[i0] == 'A' and form[i0 + 1] == 'B':
                B1 = Ss[i0 + 1][:, None, None] * B1
            elif form[i0] == 'A' and form[i0 + 1] == 'A':
                B1 = B1 * Ss[i0 + 2][None, None, :]
            theta = np.tensordot(B0, B1, axes=[2, 0])
            n_l, n_r = theta.shape[0] * d, d * theta.shape[3]
            theta = theta.reshape(n_l, n_r)
            E0, U, S, VH, env, N = bond_update(LHeff, RHeff, theta, trunc_par, lanczos_par, mr)
            chi = len(S)
            Bs[i0] = U.reshape(n_l // d, d, chi)
            Bs[i0 + 1] = VH.reshape(chi, d, n_r // d)
            form[i0], form[i0 + 1] = 'A', 'B'
            Ss[i0 + 1] = S
            Sbond[i0 + 1] = float(-np.sum(S**2 * np.log(S**2)))
            if mr:
                LP[i0 + 1] = env
            else:
                RP[i0] = env
        sweeps += 1
        E_sweeps.append(E0)
        S_sweeps.append(max(Sbond))
        if sweeps >= max(min_sweeps, 2) and len(E_sweeps) >= 2:
            dE = abs(E_sweeps[-1] - E_sweeps[-2]) / max(abs(E_sweeps[-1]), 1.)
            dS = abs(S_sweeps[-1] - S_sweeps[-2])
            if dE < max_E_err and dS < max_S_err:
                break
    return {'E': E0, 'S': Sbond[1:L], 'Ss': Ss, 'Bs': Bs, 'sweeps': sweeps, 'E_sweeps': E_sweeps}
