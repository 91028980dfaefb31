#!/usr/bin/env python
"""Golden vectors for the remaining B1 surface of the DMRG/TEBD path: Array.take_slice, add_leg, extend,
npc.concatenate -- produced by the UNMODIFIED reference on seeded inputs (build container only):

    TENPY_NO_CYTHON=1 PYTHONPATH=/root/reference python tests/golden/make_golden_b1.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402  (imports the reference, dump helpers)

npc = mg.npc


def main():
    out = {}
    rng = np.random.default_rng(271828)
    ch1 = npc.ChargeInfo([1], ['U1'])
    ch2 = npc.ChargeInfo([1, 2], ['N', 'P'])
    ch0 = npc.ChargeInfo()
    case = 0
    for chinfo in (ch1, ch2, ch0):
        legs = [mg.rand_leg(rng, chinfo, 7, +1), mg.rand_leg(rng, chinfo, 5, -1), mg.rand_leg(rng, chinfo, 6, +1),
                mg.rand_leg(rng, chinfo, 4, -1)]
        a = mg.rand_array(rng, legs, labels=['a', 'b', 'c', 'd'])
        mg.dump_array('c%d_a' % case, a, out)
        # take_slice: one axis / two axes (by label), indices chosen inside stored sectors when possible
        i1 = int(rng.integers(0, 5))
        out['c%d_ts1_idx' % case] = np.int64(i1)
        mg.dump_array('c%d_ts1' % case, a.take_slice(i1, 'b'), out)
        i2 = [int(rng.integers(0, 7)), int(rng.integers(0, 4))]
        out['c%d_ts2_idx' % case] = np.array(i2, dtype=np.int64)
        mg.dump_array('c%d_ts2' % case, a.take_slice(i2, ['a', 'd']), out)
        # add_leg: new leg in the middle and in front (the reference's add_leg does not support axis=rank)
        new_leg = mg.rand_leg(rng, chinfo, 5, -1)
        mg.dump_leg('c%d_al_leg' % case, new_leg, out)
        j = int(rng.integers(0, 5))
        out['c%d_al_idx' % case] = np.int64(j)
        mg.dump_array('c%d_al2' % case, a.add_leg(new_leg, j, axis=2, label='n'), out)
        mg.dump_array('c%d_al0' % case, a.add_leg(new_leg, j, axis=0, label='n'), out)
        # extend by a leg and by an int
        extra = mg.rand_leg(rng, chinfo, 3, +1)
        mg.dump_leg('c%d_ext_leg' % case, extra, out)
        mg.dump_array('c%d_ext' % case, a.extend('c', extra), out)
        mg.dump_array('c%d_exti' % case, a.extend('b', 2), out)
        # concatenate three arrays along axis 1 (same other legs, same qtotal; one with opposite qconj on the axis)
        b_legs = list(legs)
        b_legs[1] = mg.rand_leg(rng, chinfo, 3, -1)
        b = mg.rand_array(rng, b_legs, labels=['a', 'b', 'c', 'd'])
        c_legs = list(legs)
        c_legs[1] = mg.rand_leg(rng, chinfo, 4, +1)
        c = mg.rand_array(rng, c_legs, labels=['a', 'b', 'c', 'd'])
        mg.dump_array('c%d_b' % case, b, out)
        mg.dump_array('c%d_c' % case, c, out)
        mg.dump_array('c%d_cat' % case, npc.concatenate([a, b, c], axis='b'), out)
        # qr with the unique sign convention (pos_diag_R), default inner leg and with qtotal_Q / inner_qconj=-1
        mleg0, mleg1 = mg.rand_leg(rng, chinfo, 9, +1), mg.rand_leg(rng, chinfo, 8, -1)
        mat = mg.rand_array(rng, [mleg0, mleg1], labels=['x', 'y'])
        mg.dump_array('c%d_qr_m' % case, mat, out)
        Q, R = npc.qr(mat, inner_labels=['q', 'r'], pos_diag_R=True)
        mg.dump_array('c%d_qr_Q' % case, Q, out)
        mg.dump_array('c%d_qr_R' % case, R, out)
        tall = mat.combine_legs(['x'])          # a piped first leg
        matq = mg.rand_array(rng, [mleg0, mleg1], qtotal=mleg0.get_charge(0) + mleg1.get_charge(0), labels=['x', 'y'])
        mg.dump_array('c%d_qr_mq' % case, matq, out)
        Q, R = npc.qr(matq, inner_labels=['q', 'r'], pos_diag_R=True, qtotal_Q=matq.qtotal, inner_qconj=-1)
        mg.dump_array('c%d_qr_Qq' % case, Q, out)
        mg.dump_array('c%d_qr_Rq' % case, R, out)
        case += 1
    out['n_cases'] = np.int64(case)
    np.savez_compressed(os.path.join(HERE, 'b1_ops.npz'), **out)
    print('wrote b1_ops.npz with', len(out), 'entries')


if __name__ == '__main__':
    main()
