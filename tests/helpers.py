"""Test helpers: load the golden vectors (tests/golden/*.npz, generated from the reference by
tests/golden/make_golden.py) into oracle `OArray`s and into product `Array`s, and compare them."""
import os

import numpy as np

from oracle import npc_blocks as ob

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def load(name):
    return dict(np.load(os.path.join(GOLDEN, name), allow_pickle=False))


# ------------------------------------------------------------------ golden -> oracle
def oleg_from(g, prefix, mod, check_pipe=True):
    slices, charges, qconj = g[prefix + '_slices'], g[prefix + '_charges'], int(g[prefix + '_qconj'])
    if prefix + '_pipe_nlegs' in g:
        subs = [oleg_from(g, prefix + '_sub%d' % j, mod) for j in range(int(g[prefix + '_pipe_nlegs']))]
        leg = ob.make_pipe(subs, qconj, mod)
        if check_pipe:   # pins oracle.make_pipe against the reference's LegPipe tables
            assert np.array_equal(leg.slices, slices)
            assert np.array_equal(leg.charges, charges.reshape(leg.charges.shape))
            assert np.array_equal(leg.pipe['q_map'], g[prefix + '_pipe_qmap'])
            assert np.array_equal(leg.pipe['q_map_slices'], g[prefix + '_pipe_qmap_slices'])
        return leg
    return ob.OLeg(slices, charges, qconj)


def oarray_from(g, prefix):
    mod = g[prefix + '_mod']
    rank = int(g[prefix + '_nlegs'])
    legs = [oleg_from(g, prefix + '_leg%d' % i, mod) for i in range(rank)]
    qdata = g[prefix + '_qdata'].reshape(-1, rank)
    data = g[prefix + '_data']
    blocks, at = [], 0
    for q in qdata:
        shape = [int(l.sizes()[qi]) for l, qi in zip(legs, q)]
        n = int(np.prod(shape))
        blocks.append(data[at:at + n].reshape(shape))
        at += n
    assert at == len(data)
    labels = [str(l) if str(l) != '' else None for l in g[prefix + '_labels']]
    return ob.OArray(legs, mod, g[prefix + '_qtotal'], qdata, blocks, labels)


# ------------------------------------------------------------------ oracle -> product and back
def leg_to_product(oleg, chinfo):
    from tenpy_b200.linalg.charges import LegCharge, LegPipe
    if oleg.pipe is not None:
        subs = [leg_to_product(l, chinfo) for l in oleg.pipe['legs']]
        return LegPipe(subs, qconj=oleg.qconj)
    return LegCharge.from_qind(chinfo, oleg.slices, oleg.charges, oleg.qconj)


def to_product(oarr, labels=None):
    from tenpy_b200.linalg import np_conserved as npc
    chinfo = npc.ChargeInfo(list(oarr.mod))
    legs = [leg_to_product(l, chinfo) for l in oarr.legs]
    if labels is None:
        labels = oarr.labels
    return npc.Array.from_blocks(legs, oarr.qdata, oarr.blocks, oarr.qtotal, labels)


def to_oracle(arr):
    """product Array -> OArray (host copy of the blocks)"""
    def conv(leg):
        from tenpy_b200.linalg.charges import LegPipe
        if isinstance(leg, LegPipe):
            return ob.make_pipe([conv(l) for l in leg.legs], leg.qconj, arr.chinfo.mod)
        return ob.OLeg(leg.slices, leg.charges, leg.qconj)
    return ob.OArray([conv(l) for l in arr.legs], arr.chinfo.mod, arr.qtotal, arr._layout.qdata, arr.get_blocks_host(),
                     arr.get_leg_labels())


def assert_same_structure(x, y):
    """identical legs (slices / charges / qconj), qtotal and block table (bit-exact integer work)"""
    assert len(x.legs) == len(y.legs)
    for lx, ly in zip(x.legs, y.legs):
        assert np.array_equal(lx.slices, ly.slices)
        assert np.array_equal(lx.charges, ly.charges)
        assert lx.qconj == ly.qconj
    assert np.array_equal(x.qtotal, y.qtotal)
    xs, ys = x.sorted(), y.sorted()
    assert np.array_equal(xs.qdata, ys.qdata), (xs.qdata, ys.qdata)


def assert_close(x, y, tol=1e-13, structure=True):
    """same structure and blocks equal within `tol` (absolute, relative to the largest entry)"""
    if structure:
        assert_same_structure(x, y)
        xs, ys = x.sorted(), y.sorted()
        scale = max([1.] + [float(np.max(np.abs(b))) for b in ys.blocks if b.size])
        for bx, by in zip(xs.blocks, ys.blocks):
            assert bx.shape == by.shape
            assert np.max(np.abs(bx - by)) <= tol * scale, np.max(np.abs(bx - by))
    else:
        dx, dy = x.to_dense(), y.to_dense()
        assert np.max(np.abs(dx - dy)) <= tol * max(1., np.max(np.abs(dy)))
