"""dev probe (not a test): effective-H matvec on synthetic random-charge Arrays of the BASELINE.json configs[2]/[3]
shapes (generator modelled on the reference's tests/benchmark/tensordot_npc.py:36-51).

    python tests/dev_blocksparse_probe.py xxz   1024 12     # U(1), chi=1024, ~12 sectors, d=2, D=5
    python tests/dev_blocksparse_probe.py hub   2048 40     # U(1)xU(1), chi=2048, ~40 sectors, d=4, D=6
"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, '.')
from tenpy_b200 import backend
from tenpy_b200.linalg import np_conserved as npc
from tenpy_b200.linalg.charges import ChargeInfo, LegCharge


def random_sector_leg(rng, chinfo, ind_len, n_sectors, qconj, spread):
    """leg with `n_sectors` sectors of random sizes (random partition of ind_len), sorted charges"""
    cuts = np.sort(rng.choice(np.arange(1, ind_len), size=n_sectors - 1, replace=False))
    slices = np.concatenate(([0], cuts, [ind_len]))
    qn = chinfo.qnumber
    charges = set()
    while len(charges) < n_sectors:
        charges.add(tuple(int(x) for x in rng.integers(-spread, spread + 1, size=qn)))
    charges = np.array(sorted(charges))
    charges = charges[np.lexsort(charges.T)]
    return LegCharge.from_qind(chinfo, slices, charges, qconj)


def main():
    kind, chi, nsec = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    rng = np.random.default_rng(0)
    if kind == 'xxz':
        ci = ChargeInfo([1], ['2*Sz'])
        p = LegCharge.from_qflat(ci, [[-1], [1]], +1)
        wq = np.array([[0], [2], [-2], [0], [0]])
        spread = 16
    else:
        ci = ChargeInfo([1, 1], ['N', '2*Sz'])
        p = LegCharge.from_qflat(ci, [[0, 0], [1, -1], [1, 1], [2, 0]], +1)
        wq = np.array([[0, 0], [1, 1], [-1, -1], [1, -1], [-1, 1], [0, 0]])
        spread = 8
    D = len(wq)
    vL = random_sector_leg(rng, ci, chi, nsec, +1, spread)
    vR = random_sector_leg(rng, ci, chi, nsec, -1, spread)
    w = LegCharge.from_qind(ci, np.arange(D + 1), wq, -1)
    gen = lambda shape: rng.standard_normal(shape)
    t0 = time.time()
    L4 = npc.Array.from_func(gen, [vL, p, w, vL.conj(), p.conj()], labels=['vR*', 'p0', 'wR', 'vR', 'p0*'])
    LHeff = L4.combine_legs([['vR*', 'p0'], ['vR', 'p0*']], qconj=[+1, -1], new_axes=[0, 2])
    R4 = npc.Array.from_func(gen, [w.conj(), p.conj(), vR.conj(), p, vR], labels=['wL', 'p1*', 'vL', 'p1', 'vL*'])
    RHeff = R4.combine_legs([['p1', 'vL*'], ['p1*', 'vL']], qconj=[-1, +1], new_axes=[2, 1])
    theta = npc.Array.from_func(gen, [LHeff.get_leg('(vR.p0*)').conj(), RHeff.get_leg('(p1*.vL)').conj()],
                                labels=['(vL.p0)', '(p1.vR)'])
    print('setup %.1fs; blocks: LHeff %d, theta %d, RHeff %d; theta %s' % (time.time() - t0, LHeff.stored_blocks,
                                                                       theta.stored_blocks, RHeff.stored_blocks, theta.shape))

    def mv(th):
        t = npc.tensordot(LHeff, th, axes=['(vR.p0*)', '(vL.p0)'])
        t = npc.tensordot(t, RHeff, axes=[['wR', '(p1.vR)'], ['wL', '(p1*.vL)']])
        return t
    for _ in range(3):
        out = mv(theta)
    from tenpy_b200.linalg.np_conserved import _PLAN_CACHE
    flops = sum(v[2].flops for v in _PLAN_CACHE.values())
    ngemm = sum(v[2].n_pairs for v in _PLAN_CACHE.values())
    torch.cuda.synchronize()
    reps = 20
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    ev0.record()
    for _ in range(reps):
        out = mv(theta)
    ev1.record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / reps
    ms = ev0.elapsed_time(ev1) / reps
    dense = 4. * D * p.ind_len**3 * chi**3
    print('%s chi=%d: %d GEMMs/matvec, %.3e flop (%.2f%% of dense), %.3f ms/matvec (wall %.3f ms) -> %.1f GFLOP/s' %
          (kind, chi, ngemm, flops, 100 * flops / dense, ms, wall * 1e3, flops / ms / 1e6))
    # svd of theta (block-diagonal batch)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    U, S, VH = npc.svd(theta)
    torch.cuda.synchronize()
    print('block svd: %d blocks, largest %s, %.1f ms, sweeps %d' % (theta.stored_blocks, tuple(theta._layout.shapes.max(axis=0)),
                                                                (time.perf_counter() - t0) * 1e3, npc.svd_stats['jacobi_sweeps'][-1]))
    rec = npc.tensordot(U.scale_axis(S, 1), VH, axes=1)
    print('svd rec err', npc.norm(rec - theta) / npc.norm(theta))


if __name__ == '__main__':
    main()
