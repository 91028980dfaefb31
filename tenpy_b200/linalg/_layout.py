"""Packed HBM block layout of an Array and the host-side *index plans* for block data movement.

An :class:`~tenpy_b200.linalg.np_conserved.Array` stores all its charge blocks in ONE contiguous HBM
buffer.  :class:`BlockLayout` is the immutable table describing it:

* ``qdata``   (nblocks, rank) int64, lex-sorted (last leg = primary key, the reference's convention for
  ``Array._qdata`` with ``_qdata_sorted=True``, tenpy/linalg/np_conserved.py:1431),
* ``shapes``  (nblocks, rank) block extents (from the legs' block sizes),
* ``offsets`` element offset of every block, aligned to 16 elements (128 B); padding is kept zero so that
  BLAS-1 style kernels can stream over the whole buffer,
* ``size``    total number of elements of the buffer.

The functions ``plan_*`` compute, with vectorised integer numpy only, the copy / take / scale records
consumed by the CUDA kernels of ``csrc/move.cu`` (formats: include/b200npc.h).  They restate the block
bookkeeping of the reference's `_combine_legs_worker` (np_conserved.py:4404), `_split_legs_worker`
(:4483), `Array.itranspose` (:2057), `Array.iproject` (:1914) and `Array.iscale_axis` (:2108).
"""
# Copyright (C) 2026 tenpy_b200 authors. Apache-2.0.

import itertools

import numpy as np

from .charges import _lexsort_rows, _row_change_points

__all__ = ['BlockLayout', 'ALIGN', 'COPY_REC', 'COPY_MAXRANK', 'plan_transpose', 'plan_combine', 'plan_split',
           'plan_project', 'plan_scale_axis', 'plan_take_slice', 'plan_add_leg', 'plan_concatenate']

ALIGN = 16
COPY_REC = 22
COPY_MAXRANK = 6
_uid = itertools.count(1)


def _aligned_offsets(sizes):
    padded = (np.asarray(sizes, dtype=np.int64) + (ALIGN - 1)) // ALIGN * ALIGN
    offs = np.zeros(len(padded) + 1, dtype=np.int64)
    np.cumsum(padded, out=offs[1:])
    return offs[:-1].copy(), int(offs[-1])


def _contig_strides(shapes):
    """row-major strides (elements) for every row of `shapes` (nblocks, rank)."""
    shapes = np.asarray(shapes, dtype=np.int64)
    st = np.ones_like(shapes)
    for ax in range(shapes.shape[1] - 2, -1, -1):
        st[:, ax] = st[:, ax + 1] * shapes[:, ax + 1]
    return st


_INTERN = {}
_INTERN_MAX = 50000


class BlockLayout:
    """Immutable block table of a packed buffer (see module doc-string).

    Layouts are *interned*: constructing a layout with a block table (``qdata``, ``shapes``) that was seen
    before returns the very same object.  All caches hanging off a layout (transpose / combine / split /
    scale records in ``layout.cache``, the contraction plans of ``np_conserved.tensordot`` keyed on layout
    identity) therefore hit whenever the block *structure* repeats -- every Lanczos iteration of a bond and
    the same bond in the next sweep -- and the integer bookkeeping that the reference redoes per call
    (`_tensordot_pre_sort`, pyx:1337; `_tensordot_match_charges`, pyx:1382) is done once per structure.
    A cache entry that also depends on leg data not visible in the block table (pipes) has to carry
    ``LegCharge.content_key()`` of those legs in its key.
    """
    __slots__ = ('qdata', 'shapes', 'sizes', 'offsets', 'size', 'uid', 'rank', 'nblocks', 'cache', 'has_padding')

    def __new__(cls, qdata, shapes):
        qdata = np.ascontiguousarray(qdata, dtype=np.int64)
        shapes = np.ascontiguousarray(shapes, dtype=np.int64)
        if qdata.ndim != 2 or qdata.shape != shapes.shape:
            raise ValueError('qdata/shapes mismatch')
        key = (qdata.shape, qdata.tobytes(), shapes.tobytes())
        self = _INTERN.get(key)
        if self is not None:
            return self
        self = object.__new__(cls)
        self.qdata = qdata = qdata.copy()     # private, read-only copies: the table is shared between Arrays
        self.shapes = shapes = shapes.copy()
        qdata.flags.writeable = False
        shapes.flags.writeable = False
        self.nblocks, self.rank = qdata.shape
        self.sizes = np.prod(shapes, axis=1, dtype=np.int64) if self.rank else np.ones(self.nblocks, np.int64)
        self.offsets, self.size = _aligned_offsets(self.sizes)
        self.has_padding = bool(np.any(self.sizes % ALIGN))
        self.uid = next(_uid)
        self.cache = {}
        if len(_INTERN) >= _INTERN_MAX:
            _INTERN.clear()     # old layouts stay valid (Arrays hold them), they are just no longer canonical
        _INTERN[key] = self
        return self

    def __reduce__(self):
        return (BlockLayout, (np.array(self.qdata), np.array(self.shapes)))

    @classmethod
    def from_legs(cls, legs, qdata, presorted=False):
        """Layout for the given qindex table; sorts the rows unless `presorted`.

        Returns ``(layout, perm)`` with ``layout.qdata == qdata[perm]``."""
        qdata = np.asarray(qdata, dtype=np.int64).reshape(-1, len(legs))
        if presorted or qdata.shape[0] < 2:
            perm = np.arange(qdata.shape[0], dtype=np.intp)
        else:
            perm = _lexsort_rows(qdata)
            qdata = qdata[perm]
        shapes = np.empty_like(qdata)
        for ax, leg in enumerate(legs):
            shapes[:, ax] = leg.get_block_sizes()[qdata[:, ax]]
        return cls(qdata, shapes), perm

    def same_blocks(self, other):
        return self is other or (self.qdata.shape == other.qdata.shape and np.array_equal(self.qdata, other.qdata)
                                 and np.array_equal(self.shapes, other.shapes))

    def strides(self):
        return _contig_strides(self.shapes)


def _copy_records(soff, doff, shape_it, sstride, dstride):
    """assemble copy records (n, 22): [soff, doff, n_elem, rank, shape[6], sstride[6], dstride[6]]."""
    n, r = shape_it.shape
    if r > COPY_MAXRANK:
        raise ValueError('copy rank {0} exceeds {1}'.format(r, COPY_MAXRANK))
    rec = np.zeros((n, COPY_REC), dtype=np.int64)
    rec[:, 0] = soff
    rec[:, 1] = doff
    rec[:, 2] = np.prod(shape_it, axis=1) if r else 1
    rec[:, 3] = r
    rec[:, 4:4 + COPY_MAXRANK] = 1
    rec[:, 4:4 + r] = shape_it
    rec[:, 4 + COPY_MAXRANK:4 + COPY_MAXRANK + r] = sstride
    rec[:, 4 + 2 * COPY_MAXRANK:4 + 2 * COPY_MAXRANK + r] = dstride
    return rec


def _merge_dims(groups, shape_it, sstride, dstride):
    """merge iteration dims listed in `groups` (lists of consecutive dim indices that are contiguous in
    both source and destination); returns merged (shape, sstride, dstride)."""
    n = shape_it.shape[0]
    g = len(groups)
    sh = np.ones((n, g), dtype=np.int64)
    ss = np.ones((n, g), dtype=np.int64)
    ds = np.ones((n, g), dtype=np.int64)
    for k, grp in enumerate(groups):
        sh[:, k] = np.prod(shape_it[:, grp], axis=1)
        ss[:, k] = sstride[:, grp[-1]]
        ds[:, k] = dstride[:, grp[-1]]
    return sh, ss, ds


def plan_transpose(layout, perm):
    """Blocks of `layout` with legs permuted by `perm` (new leg j = old leg perm[j]).

    Returns ``(new_layout, records)``; the kernel writes each new block contiguously."""
    perm = list(perm)
    qd = layout.qdata[:, perm]
    order = _lexsort_rows(qd) if qd.shape[0] > 1 else np.arange(qd.shape[0], dtype=np.intp)
    new = BlockLayout(qd[order], layout.shapes[order][:, perm])
    old_st = layout.strides()[order][:, perm]          # source stride of every new dim
    new_st = new.strides()
    # runs of consecutive old legs stay contiguous in both -> merge
    groups = [[0]] if perm else []
    for j in range(1, len(perm)):
        if perm[j] == perm[j - 1] + 1:
            groups[-1].append(j)
        else:
            groups.append([j])
    sh, ss, ds = _merge_dims(groups, new.shapes, old_st, new_st)
    rec = _copy_records(layout.offsets[order], new.offsets, sh, ss, ds)
    return new, rec


def plan_combine(layout, legs_old, combine_legs, new_axes, pipes, res_legs):
    """Index plan of ``combine_legs``: old blocks -> sub-slices of fused blocks.

    `combine_legs`: list of lists of old axes; `new_axes`: positions of the pipes in the result (ascending
    order as in the reference); `pipes`: the LegPipes; `res_legs`: all legs of the result.
    Returns ``(new_layout, records)``.  The destination buffer must be zero-initialised.
    Restates np_conserved.py:4404-4478.
    """
    rank_old = layout.rank
    rank_new = len(res_legs)
    combined = [ax for cl in combine_legs for ax in cl]
    non_combined = [ax for ax in range(rank_old) if ax not in combined]
    non_new_axes = [ax for ax in range(rank_new) if ax not in new_axes]
    # order of old axes in the (virtually) transposed array
    transp = [None] * rank_new
    for na, cl in zip(new_axes, combine_legs):
        transp[na] = list(cl)
    for na, oa in zip(non_new_axes, non_combined):
        transp[na] = [oa]
    nb = layout.nblocks
    qdata = np.empty((nb, rank_new), dtype=np.int64)
    block_start = np.zeros((nb, rank_new), dtype=np.int64)
    qdata[:, non_new_axes] = layout.qdata[:, non_combined]
    for pipe, cl, na in zip(pipes, combine_legs, new_axes):
        rows = pipe._map_incoming_qind(layout.qdata[:, cl])
        qdata[:, na] = pipe.q_map[rows, 2]
        block_start[:, na] = pipe.q_map[rows, 0]
    order = _lexsort_rows(qdata) if nb > 1 else np.arange(nb, dtype=np.intp)
    qdata = qdata[order]
    block_start = block_start[order]
    diffs = _row_change_points(qdata)
    new_qdata = qdata[diffs[:-1]]
    new = BlockLayout.from_legs(res_legs, new_qdata, presorted=True)[0]
    # index of the target block for every old block
    target = np.repeat(np.arange(len(diffs) - 1), np.diff(diffs))
    big_st = new.strides()[target]                     # (nb, rank_new)
    old_shapes = layout.shapes[order]
    old_st = layout.strides()[order]
    # iteration dims = old axes in transposed order
    flat_axes = [ax for grp in transp for ax in grp]
    shape_it = old_shapes[:, flat_axes]
    sstride = old_st[:, flat_axes]
    dstride = np.empty_like(shape_it)
    col = 0
    for na, grp in enumerate(transp):
        # within a pipe the sub-block is row-major over the incoming legs
        inner = np.ones(nb, dtype=np.int64)
        for l in range(len(grp) - 1, -1, -1):
            dstride[:, col + l] = big_st[:, na] * inner
            inner = inner * old_shapes[:, grp[l]]
        col += len(grp)
    doff = new.offsets[target] + np.sum(block_start * big_st, axis=1)
    # merge dims that are adjacent old axes (contiguous in the source) and adjacent in the destination
    groups = [[0]] if flat_axes else []
    for j in range(1, len(flat_axes)):
        if flat_axes[j] == flat_axes[j - 1] + 1 and _same_dst_group(transp, j):
            groups[-1].append(j)
        else:
            groups.append([j])
    sh, ss, ds = _merge_dims(groups, shape_it, sstride, dstride)
    rec = _copy_records(layout.offsets[order], doff, sh, ss, ds)
    return new, rec


def _same_dst_group(transp, j):
    """True if flat iteration dims j-1 and j belong to the same pipe (destination contiguous)."""
    col = 0
    for grp in transp:
        if col < j < col + len(grp):
            return True
        col += len(grp)
    return False


def plan_split(layout, legs_old, split_axes, res_legs):
    """Index plan of ``split_legs`` (inverse of :func:`plan_combine`); all sub-blocks are kept, like the
    reference (np_conserved.py:4483-4570 ignores its `cutoff`).  Returns ``(new_layout, records)``."""
    rank_old = layout.rank
    nb = layout.nblocks
    pipes = [legs_old[ax] for ax in split_axes]
    # number of q_map rows per old block and pipe
    beg = np.zeros((nb, len(pipes)), dtype=np.int64)
    cnt = np.ones((nb, len(pipes)), dtype=np.int64)
    for j, (pipe, ax) in enumerate(zip(pipes, split_axes)):
        qi = layout.qdata[:, ax]
        beg[:, j] = pipe.q_map_slices[qi]
        cnt[:, j] = pipe.q_map_slices[qi + 1] - pipe.q_map_slices[qi]
    per_block = np.prod(cnt, axis=1) if len(pipes) else np.ones(nb, dtype=np.int64)
    old_idx = np.repeat(np.arange(nb), per_block)
    n_new = len(old_idx)
    # enumerate the cartesian product of q_map rows within every old block (row-major over the pipes)
    start = np.concatenate(([0], np.cumsum(per_block)))[:-1]
    local = np.arange(n_new) - np.repeat(start, per_block)
    rows = np.empty((n_new, len(pipes)), dtype=np.int64)
    rem = local
    for j in range(len(pipes) - 1, -1, -1):
        c = cnt[old_idx, j]
        rows[:, j] = beg[old_idx, j] + rem % c
        rem = rem // c
    # new axes bookkeeping
    new_axis_of = []      # for every old axis: list of new axes
    na = 0
    for ax in range(rank_old):
        if ax in split_axes:
            k = legs_old[ax].nlegs
            new_axis_of.append(list(range(na, na + k)))
            na += k
        else:
            new_axis_of.append([na])
            na += 1
    rank_new = na
    new_qdata = np.empty((n_new, rank_new), dtype=np.int64)
    src_start = np.zeros((n_new, rank_old), dtype=np.int64)
    for ax in range(rank_old):
        if ax in split_axes:
            j = list(split_axes).index(ax)
            qm = pipes[j].q_map[rows[:, j]]
            new_qdata[:, new_axis_of[ax]] = qm[:, 3:]
            src_start[:, ax] = qm[:, 0]
        else:
            new_qdata[:, new_axis_of[ax][0]] = layout.qdata[old_idx, ax]
    new, order = BlockLayout.from_legs(res_legs, new_qdata)
    old_idx = old_idx[order]
    src_start = src_start[order]
    old_st = layout.strides()[old_idx]
    soff = layout.offsets[old_idx] + np.sum(src_start * old_st, axis=1)
    # iteration dims = new axes; the source of a sub-leg inside a pipe is row-major over the sub-legs
    sstride = np.empty((n_new, rank_new), dtype=np.int64)
    for ax in range(rank_old):
        nas = new_axis_of[ax]
        inner = np.ones(n_new, dtype=np.int64)
        for l in range(len(nas) - 1, -1, -1):
            sstride[:, nas[l]] = old_st[:, ax] * inner
            inner = inner * new.shapes[:, nas[l]]
    dstride = new.strides()
    groups = []
    for ax in range(rank_old):
        groups.append(list(new_axis_of[ax]))   # sub-legs of one pipe are contiguous in source and destination
    sh, ss, ds = _merge_dims(groups, new.shapes, sstride, dstride)
    rec = _copy_records(soff, new.offsets, sh, ss, ds)
    return new, rec


def plan_project(layout, legs_old, axis, map_qind, block_masks, new_leg, res_legs):
    """Index plan of ``iproject`` along one axis.

    Returns ``(new_layout, records (n,7), index_pool)`` for ``b200_take_blocks_f64``."""
    qi_old = layout.qdata[:, axis]
    qi_new = map_qind[qi_old]
    keep = np.nonzero(qi_new >= 0)[0]
    new_qdata = layout.qdata[keep].copy()
    new_qdata[:, axis] = qi_new[keep]
    new, order = BlockLayout.from_legs(res_legs, new_qdata)
    keep = keep[order]
    # index pool: for every *new* qindex of the projected leg the kept positions within the old block
    pool_off = np.zeros(len(block_masks) + 1, dtype=np.int64)
    pool = []
    for j, bm in enumerate(block_masks):
        idx = np.nonzero(bm)[0]
        pool.append(idx)
        pool_off[j + 1] = pool_off[j] + len(idx)
    pool = np.concatenate(pool).astype(np.int64) if pool else np.zeros(0, np.int64)
    old_shapes = layout.shapes[keep]
    outer = np.prod(old_shapes[:, :axis], axis=1)
    inner = np.prod(old_shapes[:, axis + 1:], axis=1)
    rec = np.zeros((len(keep), 7), dtype=np.int64)
    rec[:, 0] = layout.offsets[keep]
    rec[:, 1] = new.offsets
    rec[:, 2] = outer
    rec[:, 3] = new.shapes[:, axis]
    rec[:, 4] = inner
    rec[:, 5] = old_shapes[:, axis]
    rec[:, 6] = pool_off[new.qdata[:, axis]]
    return new, rec, pool


def plan_scale_axis(layout, leg, axis):
    """records (n,5) for ``b200_scale_axis_f64``: [off, outer, len, inner, s_off]."""
    rec = np.zeros((layout.nblocks, 5), dtype=np.int64)
    rec[:, 0] = layout.offsets
    rec[:, 1] = np.prod(layout.shapes[:, :axis], axis=1)
    rec[:, 2] = layout.shapes[:, axis]
    rec[:, 3] = np.prod(layout.shapes[:, axis + 1:], axis=1)
    rec[:, 4] = leg.slices[layout.qdata[:, axis]]
    return rec


def plan_take_slice(layout, axes, qidx, ridx):
    """``A[..., i, ...]`` on the block level (reference `Array.take_slice`, np_conserved.py:1037): keep the blocks
    whose qindex on `axes` equals `qidx`, take index `ridx` (inside the block) there and drop these legs.

    Returns ``(new_layout, records)``."""
    axes = [int(a) for a in axes]
    keep_axes = [a for a in range(layout.rank) if a not in axes]
    keep = np.all(layout.qdata[:, axes] == np.asarray(qidx, dtype=np.int64)[None, :], axis=1)
    idx = np.nonzero(keep)[0]
    # dropping columns that are constant over the kept rows preserves the lex order
    new = BlockLayout(layout.qdata[np.ix_(idx, keep_axes)], layout.shapes[np.ix_(idx, keep_axes)])
    st = layout.strides()[idx]
    soff = layout.offsets[idx] + (st[:, axes] * np.asarray(ridx, dtype=np.int64)[None, :]).sum(axis=1)
    rec = _copy_records(soff, new.offsets, new.shapes, st[:, keep_axes], new.strides())
    return new, rec


def plan_add_leg(layout, axis, qi, ri, block_size):
    """Insert a leg before `axis` and put the data at index (`qi`, `ri`) of it (reference `Array.add_leg`,
    np_conserved.py:1130).  Returns ``(new_layout, records)``; the destination has to be zero-filled."""
    qd = np.insert(layout.qdata, axis, int(qi), axis=1)
    sh = np.insert(layout.shapes, axis, int(block_size), axis=1)
    new = BlockLayout(qd, sh)            # a constant column keeps the lex order
    nst = new.strides()
    old_axes = [a for a in range(new.rank) if a != axis]
    doff = new.offsets + int(ri) * nst[:, axis]
    rec = _copy_records(layout.offsets, doff, layout.shapes, layout.strides(), nst[:, old_axes])
    return new, rec


def plan_concatenate(layouts, legs, axis, shifts):
    """Stack block tables along `axis` (reference `concatenate`, np_conserved.py:3027): the blocks are unchanged, the
    qindex on `axis` of array j is shifted by ``shifts[j]``; the result table is lex-sorted.

    Returns ``(new_layout, [records of array 0, records of array 1, ...])`` (flat 1-D copies)."""
    qd = []
    for lay, sh in zip(layouts, shifts):
        q = lay.qdata.copy()
        q[:, axis] += int(sh)
        qd.append(q)
    qd = np.concatenate(qd, axis=0)
    new, perm = BlockLayout.from_legs(legs, qd)
    inv = np.empty(len(perm), dtype=np.int64)
    inv[perm] = np.arange(len(perm))
    recs, at = [], 0
    for lay in layouts:
        tgt = inv[at:at + lay.nblocks]
        at += lay.nblocks
        recs.append(_copy_records(lay.offsets, new.offsets[tgt], lay.sizes[:, None], np.ones((lay.nblocks, 1), np.int64),
                                  np.ones((lay.nblocks, 1), np.int64)))
    return new, recs
