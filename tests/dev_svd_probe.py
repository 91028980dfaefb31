"""dev probe (not a test): time/convergence of the block SVD kernel on random / graded / rank-deficient matrices"""
import sys, time, numpy as np, torch
sys.path.insert(0, '.')
from tenpy_b200 import backend
lib = backend.get_lib()
rng = np.random.default_rng(0)
kind = sys.argv[1]
for n in [int(x) for x in sys.argv[2:]]:
    if kind == 'gauss':
        A = rng.standard_normal((n, n))
    elif kind == 'lowrank':
        A = rng.standard_normal((n, n // 2)) @ rng.standard_normal((n // 2, n)) / n
    elif kind == 'graded':
        q1, _ = np.linalg.qr(rng.standard_normal((n, n))); q2, _ = np.linalg.qr(rng.standard_normal((n, n)))
        A = (q1 * np.logspace(0, -20, n)) @ q2
    elif kind == 'dmrg':
        q1, _ = np.linalg.qr(rng.standard_normal((n, n)))
        S = np.exp(-np.arange(n) / (n / 40.)); S /= np.linalg.norm(S)
        A = S[:, None] * q1
        h1 = rng.standard_normal((n, n)); h1 = (h1 + h1.T) / np.sqrt(n)
        h2 = rng.standard_normal((n, n)); h2 = (h2 + h2.T) / np.sqrt(n)
        A = A + 0.05 * (h1 * S[None, :] * S[:, None]) @ A @ h2
    from tenpy_b200.linalg import np_conserved as npc
    a = npc.Array.from_ndarray_trivial(A)
    lib.profile = {}
    torch.cuda.synchronize(); t0 = time.time()
    try:
        Ua, S, Va = npc.svd(a)
    except Exception as e:
        print(n, 'FAILED', e); continue
    torch.cuda.synchronize(); dt = time.time()-t0
    print('   families', {k: (v[0], round(v[1], 2)) for k, v in lib.profile_summary().items()}); lib.profile = None
    U = Ua.to_ndarray(); V = Va.to_ndarray()
    Sref = np.linalg.svd(A, compute_uv=False)
    print(kind, n, 'sweeps', npc.svd_stats['jacobi_sweeps'][-1], 'time %.1f ms' % (dt*1e3), 'dS', np.abs(S-Sref).max(), 'rec', np.abs(U@np.diag(S)@V-A).max(),
          'orthU', np.abs(U.T@U-np.eye(n)).max(), 'orthV', np.abs(V@V.T-np.eye(n)).max(), 'completions', npc.svd_stats.get('completions'), flush=True)
