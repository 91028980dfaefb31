"""Exchange of states with a stock TeNPy installation (checkpoints written from device
Arrays resume in the reference).

The reference stores an Array as ``legs`` + ``_qdata`` (block table, intp) + ``_data`` (list of C-contiguous ndarrays),
doc/intro/npc.rst:556-620; a packed device Array holds the same block table (``BlockLayout.qdata``, lex-sorted like the
reference's ``_qdata_sorted=True``) and the blocks back to back in one HBM buffer.  `to_reference` / `from_reference`
convert between the two (one D2H / H2D copy of the buffer), `mps_to_reference` / `mps_from_reference` do the same for a
finite MPS (tensors, Schmidt values, canonical forms), so that ``pickle.dump(mps_to_reference(psi), f)`` written after a
DMRG run on the GPU is a file the unmodified reference loads and continues from -- and vice versa.

``tenpy`` (the reference) is imported lazily and only here; nothing on the compute path depends on it.
"""
# Copyright (C) 2026 tenpy_b200 authors. Apache-2.0.

import numpy as np

from ..linalg import np_conserved as npc
from ..linalg.charges import ChargeInfo, LegCharge, LegPipe

__all__ = ['to_reference', 'from_reference', 'mps_to_reference', 'mps_from_reference']


def _ref():
    import tenpy.linalg.np_conserved as rnpc
    import tenpy.linalg.charges as rch
    return rnpc, rch


def _leg_to_reference(leg, rchinfo, rch):
    if isinstance(leg, LegPipe):
        subs = [_leg_to_reference(l, rchinfo, rch) for l in leg.legs]
        pipe = rch.LegPipe(subs, qconj=leg.qconj)
        if not (np.array_equal(pipe.slices, leg.slices) and np.array_equal(pipe.charges, leg.charges)):
            raise ValueError('pipe tables differ from the reference construction')
        return pipe
    res = rch.LegCharge.from_qind(rchinfo, np.asarray(leg.slices), np.asarray(leg.charges), leg.qconj)
    return res


def _leg_from_reference(rleg, chinfo, rch):
    if isinstance(rleg, rch.LegPipe):
        subs = [_leg_from_reference(l, chinfo, rch) for l in rleg.legs]
        pipe = LegPipe(subs, qconj=rleg.qconj)
        if not (np.array_equal(pipe.slices, rleg.slices) and np.array_equal(pipe.charges, rleg.charges)):
            raise ValueError('reference pipe was not built with sort=True, bunch=True')
        return pipe
    return LegCharge.from_qind(chinfo, np.asarray(rleg.slices), np.asarray(rleg.charges), rleg.qconj)


def to_reference(arr, rchinfo=None):
    """Device :class:`~tenpy_b200.linalg.np_conserved.Array` -> ``tenpy.linalg.np_conserved.Array`` (host)."""
    rnpc, rch = _ref()
    if rchinfo is None:
        rchinfo = rch.ChargeInfo(list(arr.chinfo.mod), list(arr.chinfo.names))
    legs = [_leg_to_reference(l, rchinfo, rch) for l in arr.legs]
    res = rnpc.Array(legs, np.float64, qtotal=np.asarray(arr.qtotal), labels=arr.get_leg_labels())
    res._data = arr.get_blocks_host()
    res._qdata = np.array(arr._layout.qdata, dtype=np.intp, order='C').reshape(-1, arr.rank)
    res._qdata_sorted = True
    res.test_sanity()
    return res


def from_reference(rarr, chinfo=None):
    """``tenpy.linalg.np_conserved.Array`` (real) -> device Array (one H2D copy of the packed blocks)."""
    rnpc, rch = _ref()
    if np.iscomplexobj(np.zeros(1, dtype=rarr.dtype)):
        raise NotImplementedError('complex Arrays')
    if chinfo is None:
        chinfo = ChargeInfo(list(rarr.chinfo.mod), list(rarr.chinfo.names))
    legs = [_leg_from_reference(l, chinfo, rch) for l in rarr.legs]
    blocks = [np.asarray(b, dtype=np.float64) for b in rarr._data]
    return npc.Array.from_blocks(legs, np.asarray(rarr._qdata, dtype=np.int64).reshape(-1, rarr.rank), blocks,
                                 np.asarray(rarr.qtotal), rarr.get_leg_labels())


def mps_to_reference(psi, ref_sites):
    """Finite device MPS -> ``tenpy.networks.mps.MPS`` on the reference's `ref_sites` (e.g. ``model.lat.mps_sites()``);
    tensors, Schmidt values (1-D only) and canonical forms are carried over."""
    from tenpy.networks.mps import MPS as RMPS
    ref_sites = list(ref_sites)
    if len(ref_sites) != psi.L:
        raise ValueError('need one reference site per MPS site')
    rchinfo = ref_sites[0].leg.chinfo
    Bs = [to_reference(psi.get_B(i, form=None), rchinfo) for i in range(psi.L)]
    Ss = []
    for s in psi._S:
        if isinstance(s, npc.Array):
            raise ValueError('2-D bond matrix (a mixer is still active): call engine.mixer_cleanup() first')
        Ss.append(np.array(s, dtype=np.float64))
    res = RMPS(ref_sites, Bs, Ss, bc='finite', form=[tuple(f) for f in psi.form], norm=psi.norm)
    res.test_sanity()
    return res


def mps_from_reference(rpsi, sites):
    """``tenpy.networks.mps.MPS`` (finite, real) -> device MPS on this package's `sites`."""
    from ..networks.mps import MPS
    if rpsi.bc != 'finite':
        raise NotImplementedError('only finite MPS')
    sites = list(sites)
    chinfo = sites[0].leg.chinfo
    Bs = [from_reference(rpsi.get_B(i, form=None), chinfo) for i in range(rpsi.L)]
    Ss = [np.array(rpsi.get_SL(i), dtype=np.float64) for i in range(rpsi.L)] + [np.array(rpsi.get_SR(rpsi.L - 1))]
    return MPS(sites, Bs, Ss, 'finite', [tuple(f) for f in rpsi.form], norm=rpsi.norm)
