"""dev probe (not a test): run the two dominant kernels alone at the config-2 centre-bond shapes (for ncu)"""
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
from tenpy_b200 import backend
from tenpy_b200.linalg import np_conserved as npc

lib = backend.get_lib()
chi, d, D = 1024, 2, 3
n = chi * d
ci = npc.ChargeInfo()
lL, lR, lW = npc.LegCharge.from_trivial(n, ci, +1), npc.LegCharge.from_trivial(n, ci, -1), npc.LegCharge.from_trivial(D, ci, -1)


def rnd(legs):
    t = torch.randn(int(np.prod([l.ind_len for l in legs])), dtype=torch.float64, device='cuda')
    return npc.Array.from_device_buffer(legs, np.zeros((1, len(legs)), np.int64), t)


if sys.argv[1] == 'gemm':
    LHeff, theta, RHeff = rnd([lL, lW, lL.conj()]), rnd([lL, lR]), rnd([lW.conj(), lR.conj(), lR])
    for _ in range(6):
        t1 = npc.tensordot(LHeff, theta, axes=[2, 0])
        t2 = npc.tensordot(t1, RHeff, axes=[[1, 2], [0, 1]])
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(10):
        t1 = npc.tensordot(LHeff, theta, axes=[2, 0])
        t2 = npc.tensordot(t1, RHeff, axes=[[1, 2], [0, 1]])
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / 10
    print('matvec %.3f ms  %.2f TFLOP/s' % (ms, 4 * D * d**3 * chi**3 / ms / 1e9))
else:
    q, _ = torch.linalg.qr(torch.randn(n, n, dtype=torch.float64, device='cuda'))
    q2, _ = torch.linalg.qr(torch.randn(n, n, dtype=torch.float64, device='cuda'))
    s = torch.exp(-torch.arange(n, dtype=torch.float64, device='cuda') / 40.)
    th = npc.Array.from_device_buffer([lL, lR], np.zeros((1, 2), np.int64), ((q * s[None, :]) @ q2).reshape(-1).contiguous())
    U, S, VH = npc.svd(th)
    torch.cuda.synchronize()
    print('sweeps', npc.svd_stats['jacobi_sweeps'])
