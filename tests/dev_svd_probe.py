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
    dA = backend.to_device(A.ravel())
    dU, dS, dV = backend.zeros(n*n), backend.zeros(n), backend.zeros(n*n)
    torch.cuda.synchronize(); t0 = time.time()
    try:
        info = lib.block_svd([n],[n],[0],[0],[0],[0], dA, dU, dS, dV)
    except Exception as e:
        print(n, 'FAILED', e); continue
    torch.cuda.synchronize(); dt = time.time()-t0
    S = backend.to_host(dS); U = backend.to_host(dU).reshape(n,n); V = backend.to_host(dV).reshape(n,n)
    Sref = np.linalg.svd(A, compute_uv=False)
    k = int(np.sum(Sref > 1e-13 * Sref[0]))
    print(kind, n, 'sweeps', info[0], 'time %.1f ms' % (dt*1e3), 'dS', np.abs(S-Sref).max(), 'rec', np.abs(U@np.diag(S)@V-A).max(),
          'orthU(k=%d)' % k, np.abs(U[:, :k].T@U[:, :k]-np.eye(k)).max(), 'orthV', np.abs(V@V.T-np.eye(n)).max(), flush=True)
