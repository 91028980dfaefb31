"""Charge bookkeeping for block-sparse tensors (host side, integers only).

Host-side mirror of the reference interface ``tenpy/linalg/charges.py``
(`ChargeInfo` :39, `LegCharge` :552, `LegPipe` :1444): same class and method names, same argument
meaning, same resulting `charges` / `slices` / `q_map` tables, so that block layouts produced here are
identical to the reference's.  Everything in this module is small int64 work that stays on the host;
it feeds the contraction / reshape *plans* which the CUDA kernels execute.

Conventions (identical to the reference):

* ``charges`` is a 2D int64 array ``(block_number, qnumber)``, ``slices`` a 1D intp array of length
  ``block_number + 1``; block ``qi`` of the leg covers indices ``slices[qi]:slices[qi+1]``.
* ``qconj = +1`` means charges point inward, ``-1`` outward.  The charge rule of an Array is
  ``sum_legs qconj * charges[qindex] == qtotal  (mod `mod`)``.
* ``mod == 1`` denotes a U(1) charge, ``mod == N > 1`` a Z_N charge.
* all "lexsort"s follow ``np.lexsort``: the *last* column is the primary key.
"""
# Copyright (C) 2026 tenpy_b200 authors. Apache-2.0.

import numpy as np

__all__ = ['QTYPE', 'ChargeInfo', 'DipolarChargeInfo', 'LegCharge', 'LegPipe']

QTYPE = np.int64  # reference: charges.py:35


def _lexsort_rows(a):
    """argsort of the rows of 2D `a`, last column = primary key (np.lexsort convention)."""
    a = np.asarray(a)
    if a.shape[0] == 0 or a.shape[1] == 0:
        return np.arange(a.shape[0], dtype=np.intp)
    return np.lexsort(a.T).astype(np.intp, copy=False)


def _inverse_permutation(perm):
    inv = np.empty(len(perm), dtype=np.intp)
    inv[perm] = np.arange(len(perm), dtype=np.intp)
    return inv


def _row_change_points(rows):
    """Indices ``i`` where row ``i`` differs from row ``i-1``, including ``0`` and ``len(rows)``.

    Reference: ``charges._find_row_differences`` (charges.py:1922 / _npc_helper.pyx:635).
    """
    rows = np.asarray(rows)
    n = rows.shape[0]
    if n == 0:
        return np.zeros(1, dtype=np.intp)
    if rows.ndim == 1:
        diff = rows[1:] != rows[:-1]
    else:
        diff = np.any(rows[1:] != rows[:-1], axis=1)
    return np.concatenate(([0], np.nonzero(diff)[0] + 1, [n])).astype(np.intp)


_find_row_differences = _row_change_points


class ChargeInfo:
    """Meta-data of the conserved charges: how many, and their modulus (reference charges.py:39)."""

    trivial_shift = True     # translations do not change the charges (reference charges.py:82; False only for dipoles)

    def shift_charges(self, charges, dx):
        """charges after a translation by `dx` sites: unchanged without dipole conservation (reference charges.py:309)"""
        return charges

    def shift_charges_horizontal(self, charges, dx):
        return charges

    def __init__(self, mod=(), names=None):
        self._mod = np.array(mod, dtype=QTYPE).reshape(-1)
        self._qnumber = len(self._mod)
        self._mask = self._mod != 1  # where a modulo has to be taken
        self._mod_masked = self._mod[self._mask]
        if names is None:
            names = [''] * self._qnumber
        self.names = [str(n) for n in names]
        self.test_sanity()

    def test_sanity(self):
        if len(self.names) != self._qnumber:
            raise ValueError('names has incompatible length with mod')
        if np.any(self._mod <= 0):
            raise ValueError('mod should be > 0')

    @property
    def qnumber(self):
        return self._qnumber

    @property
    def mod(self):
        return self._mod

    @classmethod
    def add(cls, chinfos):
        """Concatenate several ChargeInfo (reference charges.py:170)."""
        mod = np.concatenate([c.mod for c in chinfos]) if len(chinfos) else []
        names = sum([c.names for c in chinfos], [])
        return cls(mod, names)

    def _charge_index(self, charge):
        if isinstance(charge, str):
            return self.names.index(charge)
        return int(charge)

    @classmethod
    def drop(cls, chinfo, charge=None):
        """`chinfo` without the given charge(s) (index / name / list; ``None`` drops all; reference charges.py:187)"""
        if charge is None:
            return cls()
        drop = [chinfo._charge_index(c) for c in (charge if isinstance(charge, (list, tuple)) else [charge])]
        keep = [i for i in range(chinfo.qnumber) if i not in drop]
        return cls([chinfo.mod[i] for i in keep], [chinfo.names[i] for i in keep])

    @classmethod
    def change(cls, chinfo, charge, new_qmod, new_name=''):
        """`chinfo` with another modulus for one charge (reference charges.py:215)"""
        i = chinfo._charge_index(charge)
        mod, names = list(chinfo.mod), list(chinfo.names)
        mod[i], names[i] = new_qmod, new_name
        return cls(mod, names)

    def make_valid(self, charges=None):
        """Take charges modulo `mod`; ``None`` gives the zero charge (reference charges.py:267)."""
        if charges is None:
            return np.zeros((self._qnumber,), dtype=QTYPE)
        charges = np.array(charges, dtype=QTYPE)  # copy
        if self._mod_masked.size:
            charges[..., self._mask] = np.mod(charges[..., self._mask], self._mod_masked)
        return charges

    def check_valid(self, charges):
        """True iff all `charges` are already reduced modulo `mod` (reference charges.py:289)."""
        charges = np.asarray(charges, dtype=QTYPE)[..., self._mask]
        return bool(np.all(np.logical_and(0 <= charges, charges < self._mod_masked)))

    def __eq__(self, other):
        if self is other:
            return True
        if not isinstance(other, ChargeInfo):
            return NotImplemented
        return self._qnumber == other._qnumber and np.array_equal(self._mod, other._mod) \
            and self.names == other.names

    def __ne__(self, other):
        r = self.__eq__(other)
        return r if r is NotImplemented else not r

    def __hash__(self):
        return hash((self._qnumber, self._mod.tobytes()))

    def __repr__(self):
        return 'ChargeInfo({0!s}, {1!s})'.format(list(self.mod), self.names)

    def __getstate__(self):
        return (self._qnumber, self._mod, self.names)

    def __setstate__(self, state):
        qnumber, mod, names = state
        self._mod = np.array(mod, dtype=QTYPE)
        self._qnumber = int(qnumber)
        self._mask = self._mod != 1
        self._mod_masked = self._mod[self._mask]
        self.names = list(names)


class DipolarChargeInfo(ChargeInfo):
    """Charges with dipole conservation (reference charges.py:331) are not provided by the B200 engine."""

    def __init__(self, *args, **kwargs):
        raise NotImplementedError('DipolarChargeInfo is outside the scope of tenpy_b200')


class LegCharge:
    """Charge data of one leg of an Array (reference charges.py:552)."""

    def __init__(self, chargeinfo, slices, charges, qconj=1):
        self.chinfo = chargeinfo
        self.slices = np.array(slices, dtype=np.intp)
        self.charges = np.array(charges, dtype=QTYPE).reshape(len(self.slices) - 1, chargeinfo.qnumber)
        self.qconj = int(qconj)
        self.sorted = False
        self.bunched = False
        self.ind_len = int(self.slices[-1])
        self.block_number = int(self.charges.shape[0])
        self._layout_key = None

    def copy(self):
        """Shallow copy (charges/slices are treated as immutable)."""
        res = LegCharge.__new__(LegCharge)
        res.__dict__.update(self.__dict__)
        res._layout_key = None
        return res

    def content_key(self):
        """Hashable value identifying the charge data of this leg (slices, charges, mod, qconj; for pipes also
        the incoming legs and `q_map`).  Used in the keys of the index-plan caches hanging off interned
        :class:`~tenpy_b200.linalg._layout.BlockLayout` objects; the qconj-independent part is cached."""
        base = self._layout_key
        if base is None:
            base = self._layout_key = self._content_base()
        return (self.qconj, base)

    def _content_base(self):
        return (self.slices.tobytes(), self.charges.tobytes(), self.chinfo.mod.tobytes())

    # --- alternative constructors (reference charges.py:758-841)
    @classmethod
    def from_trivial(cls, ind_len, chargeinfo=None, qconj=1):
        if chargeinfo is None:
            chargeinfo = ChargeInfo()
        res = cls(chargeinfo, [0, ind_len], np.zeros((1, chargeinfo.qnumber), QTYPE), qconj)
        res.sorted = res.bunched = True
        return res

    @classmethod
    def from_qflat(cls, chargeinfo, qflat, qconj=1):
        """From one charge per index; consecutive equal charges form one block."""
        qflat = np.array(qflat, dtype=QTYPE)
        ind_len = qflat.shape[0]
        qflat = qflat.reshape(ind_len, chargeinfo.qnumber)
        qflat = chargeinfo.make_valid(qflat)
        slices = _row_change_points(qflat)
        res = cls(chargeinfo, slices, qflat[slices[:-1]], qconj)
        res.sorted = res.is_sorted()
        res.bunched = res.is_bunched()
        return res

    @classmethod
    def from_qind(cls, chargeinfo, slices, charges, qconj=1):
        """Like the constructor, but makes charges valid and determines `sorted`/`bunched`."""
        charges = chargeinfo.make_valid(np.array(charges, dtype=QTYPE).reshape(len(slices) - 1, -1))
        res = cls(chargeinfo, slices, charges, qconj)
        res.sorted = res.is_sorted()
        res.bunched = res.is_bunched()
        return res

    @classmethod
    def from_add_charge(cls, legs, chargeinfo=None):
        """several legs of equal length and qconj -> one leg carrying all their charges (reference charges.py:843)"""
        legs = list(legs)
        if chargeinfo is None:
            chargeinfo = ChargeInfo.add([l.chinfo for l in legs])
        if any(l.ind_len != legs[0].ind_len or l.qconj != legs[0].qconj for l in legs):
            raise ValueError('legs to be combined need the same length and qconj')
        return cls.from_qflat(chargeinfo, np.concatenate([l.to_qflat() for l in legs], axis=1), legs[0].qconj)

    @classmethod
    def from_drop_charge(cls, leg, charge=None, chargeinfo=None):
        """`leg` without the given charge(s) (reference charges.py:875)"""
        if chargeinfo is None:
            chargeinfo = ChargeInfo.drop(leg.chinfo, charge)
        if charge is None:
            keep = []
        else:
            drop = [leg.chinfo._charge_index(c) for c in (charge if isinstance(charge, (list, tuple)) else [charge])]
            keep = [i for i in range(leg.chinfo.qnumber) if i not in drop]
        return cls.from_qflat(chargeinfo, leg.to_qflat()[:, keep], leg.qconj)

    @classmethod
    def from_change_charge(cls, leg, charge, new_qmod, new_name='', chargeinfo=None):
        """`leg` with another modulus for one charge (reference charges.py:905)"""
        if chargeinfo is None:
            chargeinfo = ChargeInfo.change(leg.chinfo, charge, new_qmod, new_name)
        return cls.from_qflat(chargeinfo, leg.to_qflat(), leg.qconj)

    @classmethod
    def from_qdict(cls, chargeinfo, qdict, qconj=1):
        """From a dict ``{charge tuple: slice}``."""
        items = sorted(((sl.start, sl.stop, q) for q, sl in qdict.items()))
        slices = [it[0] for it in items] + [items[-1][1]]
        charges = [it[2] for it in items]
        return cls.from_qind(chargeinfo, slices, charges, qconj)

    def test_sanity(self):
        sl = self.slices
        if sl.shape != (self.block_number + 1,) or self.charges.shape != (self.block_number, self.chinfo.qnumber):
            raise ValueError('wrong shapes of slices/charges')
        if sl[0] != 0 or np.any(sl[1:] <= sl[:-1]) and self.ind_len > 0:
            raise ValueError('slices have to be strictly increasing, starting at 0')
        if not self.chinfo.check_valid(self.charges):
            raise ValueError('charges invalid for ' + repr(self.chinfo))
        if self.qconj not in (-1, 1):
            raise ValueError('qconj has invalid value')

    def conj(self):
        """Shallow copy with opposite ``qconj`` (reference charges.py:979)."""
        res = self.copy()
        res.qconj = -self.qconj
        res._layout_key = self._layout_key      # same arrays: the qconj-independent part of the key carries over
        return res

    def flip_charges_qconj(self):
        """Copy with both charges and qconj negated: physically equivalent (reference :993)."""
        res = self.copy()
        res.qconj = -self.qconj
        res.charges = self.chinfo.make_valid(-self.charges)
        return res

    def to_qflat(self):
        """One charge per index, shape (ind_len, qnumber)."""
        return np.repeat(self.charges, self.get_block_sizes(), axis=0)

    def to_qdict(self):
        res = {}
        for qi in range(self.block_number):
            res[tuple(int(c) for c in self.charges[qi])] = slice(int(self.slices[qi]), int(self.slices[qi + 1]))
        if len(res) != self.block_number:
            raise ValueError('can not convert a non-blocked leg to a dict')
        return res

    def is_blocked(self):
        """True iff every charge appears in exactly one block."""
        if self.sorted and self.bunched:
            return True
        s = {tuple(c) for c in self.charges}
        return len(s) == self.block_number

    def is_sorted(self):
        if self.chinfo.qnumber == 0 or self.block_number <= 1:
            return True
        perm = _lexsort_rows(self.charges)
        return bool(np.all(perm == np.arange(len(perm))))

    def is_bunched(self):
        return len(_row_change_points(self.charges)) == self.block_number + 1

    def test_contractible(self, other):
        """Raise ValueError unless `self` can be contracted with `other` (reference :1071)."""
        if self.chinfo != other.chinfo:
            raise ValueError('incompatible ChargeInfo')
        if self.qconj != -other.qconj:
            raise ValueError('incompatible LegCharge: qconj')
        if self.ind_len != other.ind_len:
            raise ValueError('incompatible LegCharge: different ind_len')
        if self.charges is other.charges and self.slices is other.slices:
            return
        if not np.array_equal(self.slices, other.slices) or not np.array_equal(self.charges, other.charges):
            raise ValueError('incompatible LegCharge: different charges/slices\n{0!s}\nvs\n{1!s}'.format(
                self, other))

    def test_equal(self, other):
        """Raise ValueError unless the legs are equal including qconj (reference :1114)."""
        if self.chinfo != other.chinfo:
            raise ValueError('incompatible ChargeInfo')
        if self.ind_len != other.ind_len:
            raise ValueError('different ind_len')
        if self.charges is other.charges and self.slices is other.slices and self.qconj == other.qconj:
            return
        if not np.array_equal(self.slices, other.slices) or \
                not np.array_equal(self.charges * self.qconj, other.charges * other.qconj):
            raise ValueError('incompatible LegCharge: different charges')

    def __eq__(self, other):
        if not isinstance(other, LegCharge):
            return NotImplemented
        try:
            self.test_equal(other)
        except ValueError:
            return False
        return True

    def __ne__(self, other):
        r = self.__eq__(other)
        return r if r is NotImplemented else not r

    __hash__ = object.__hash__

    def get_block_sizes(self):
        return self.slices[1:] - self.slices[:-1]

    def get_slice(self, qindex):
        return slice(int(self.slices[qindex]), int(self.slices[qindex + 1]))

    def get_qindex(self, flat_index):
        """Return ``(qindex, index_within_block)`` of a flat index (reference :1172)."""
        if flat_index < 0:
            flat_index += self.ind_len
        if not 0 <= flat_index < self.ind_len:
            raise IndexError('flat index {0:d} out of bounds'.format(flat_index))
        qi = int(np.searchsorted(self.slices, flat_index, side='right')) - 1
        return qi, int(flat_index - self.slices[qi])

    def get_qindex_of_charges(self, charges):
        """qindex of the (unique) block with the given charges; requires a blocked leg."""
        charges = self.chinfo.make_valid(charges)
        match = np.nonzero(np.all(self.charges == charges, axis=1))[0]
        if len(match) != 1:
            raise ValueError('charges not found or leg not blocked')
        return int(match[0])

    def get_charge(self, qindex):
        """``charges[qindex] * qconj``."""
        return self.charges[qindex] * self.qconj

    def sort(self, bunch=True):
        """Sort (and optionally bunch) by charge.  Returns ``(perm_qind, new_leg)`` (reference :1237)."""
        if self.sorted and ((not bunch) or self.bunched):
            return np.arange(self.block_number, dtype=np.intp), self
        perm = _lexsort_rows(self.charges)
        res = self.copy()
        res.charges = self.charges[perm]
        sizes = self.get_block_sizes()[perm]
        res.slices = np.concatenate(([0], np.cumsum(sizes))).astype(np.intp)
        res.sorted = True
        res.bunched = res.is_bunched()
        if bunch and not res.bunched:
            _, res = res.bunch()
        return perm, res

    def bunch(self):
        """Merge neighbouring blocks of equal charge.  Returns ``(idx, new_leg)`` (reference :1278)."""
        if self.bunched:
            return np.arange(self.block_number + 1, dtype=np.intp), self
        idx = _row_change_points(self.charges)
        res = self.copy()
        res.charges = self.charges[idx[:-1]]
        res.slices = self.slices[idx]
        res.block_number = len(idx) - 1
        res.bunched = True
        return idx, res

    def project(self, mask):
        """Keep only the indices selected by the bool `mask`.

        Returns ``(map_qind, block_masks, projected_leg)`` as the reference (charges.py:1304).
        """
        mask = np.asarray(mask, dtype=np.bool_)
        res = self.copy()
        block_masks = [mask[b:e] for b, e in zip(self.slices[:-1], self.slices[1:])]
        new_sizes = np.array([int(np.sum(bm)) for bm in block_masks], dtype=np.intp)
        keep = np.nonzero(new_sizes)[0]
        block_masks = [block_masks[i] for i in keep]
        res.charges = self.charges[keep]
        res.slices = np.concatenate(([0], np.cumsum(new_sizes[keep]))).astype(np.intp)
        res.block_number = len(keep)
        res.ind_len = int(res.slices[-1])
        map_qind = np.full(self.block_number, -1, dtype=np.intp)
        map_qind[keep] = np.arange(len(keep), dtype=np.intp)
        return map_qind, block_masks, res

    def extend(self, extra):
        """New LegCharge with the blocks of `extra` (a LegCharge, or an int = one block of zero charge) appended
        (reference charges.py:1336)."""
        if not isinstance(extra, LegCharge):
            extra = LegCharge.from_trivial(extra, self.chinfo, self.qconj)
        bn = self.block_number
        new_slices = np.zeros(bn + extra.block_number + 1, np.intp)
        new_slices[:bn + 1] = self.slices
        new_slices[bn:] = extra.slices + self.ind_len
        new_charges = np.zeros((bn + extra.block_number, self.chinfo.qnumber), dtype=QTYPE)
        new_charges[:bn] = self.charges
        new_charges[bn:] = extra.charges if self.qconj == extra.qconj else self.chinfo.make_valid(-extra.charges)
        return LegCharge(self.chinfo, new_slices, new_charges, qconj=self.qconj)

    def charge_sectors(self):
        """Unique charge rows."""
        return np.unique(self.charges, axis=0)

    def _set_charges(self, charges):
        self.charges = charges
        self.block_number = charges.shape[0]
        self._layout_key = None

    def _set_slices(self, slices):
        self.slices = slices
        self.ind_len = int(slices[-1])
        self._layout_key = None

    def _set_block_sizes(self, block_sizes):
        self._set_slices(np.concatenate(([0], np.cumsum(block_sizes))).astype(np.intp))

    def perm_flat_from_perm_qind(self, perm_qind):
        """Translate a permutation of qindices into a permutation of flat indices."""
        begend = np.stack([self.slices[:-1], self.slices[1:]], axis=0).T
        res = [np.arange(b, e) for b, e in begend[perm_qind]]
        return np.concatenate(res).astype(np.intp) if res else np.zeros(0, np.intp)

    def __str__(self):
        return ' {0:+d}\n'.format(self.qconj) + '\n'.join(
            '{0:4d} {1!s}'.format(int(s), c) for s, c in zip(self.slices, self.charges)) + \
            '\n{0:4d}'.format(int(self.slices[-1]))

    def __repr__(self):
        return 'LegCharge({0!r}, qconj={1:+d},\n{2!r}, {3!r})'.format(self.chinfo, self.qconj,
                                                                    self.slices, self.charges)

    def __getstate__(self):
        d = dict(self.__dict__)
        d.pop('_layout_key', None)
        return d

    def __setstate__(self, state):
        self.__dict__.update(state)
        self._layout_key = None


class LegPipe(LegCharge):
    """A pipe fusing several incoming legs into one outgoing leg (reference charges.py:1444).

    Attributes as in the reference: `legs`, `nlegs`, `subshape`, `subqshape`, `q_map` with rows
    ``[b_j, b_{j+1}, I_s, i_1, ..., i_nlegs]`` (lex-sorted by ``I_s`` then ``i``), `q_map_slices`,
    `_perm`, `_strides`.  The fusion rule is
    ``pipe.charges[I] * pipe.qconj == sum_l legs[l].charges[i_l] * legs[l].qconj  (mod)``.
    """

    def __init__(self, legs, qconj=1, sort=True, bunch=True):
        chinfo = legs[0].chinfo
        LegCharge.__init__(self, chinfo, [0, 1], [[0] * chinfo.qnumber], qconj)
        self.legs = legs = tuple(legs)
        self.nlegs = len(legs)
        self.subshape = tuple(l.ind_len for l in legs)
        self.subqshape = tuple(l.block_number for l in legs)
        self.q_map = None
        self.q_map_slices = None
        self._init_from_legs(sort, bunch)

    def copy(self):
        res = LegPipe.__new__(LegPipe)
        res.__dict__.update(self.__dict__)
        res._layout_key = None
        return res

    def _content_base(self):
        # incoming legs enter with their qconj *relative* to the pipe's: invariant under conj(), which flips both
        return (LegCharge._content_base(self), self.q_map.tobytes(),
                tuple((l.qconj * self.qconj, l.content_key()[1]) for l in self.legs))

    def to_LegCharge(self):
        """Forget the incoming legs."""
        res = LegCharge.__new__(LegCharge)
        for k in ('chinfo', 'slices', 'charges', 'qconj', 'sorted', 'bunched', 'ind_len', 'block_number'):
            res.__dict__[k] = self.__dict__[k]
        res._layout_key = None
        return res

    def conj(self):
        """Shallow copy with opposite qconj; the incoming legs are conjugated as well."""
        res = LegCharge.conj(self)
        res.legs = tuple(l.conj() for l in self.legs)
        return res

    def outer_conj(self):
        """Like :meth:`conj`, but leave the incoming legs untouched (reference :1690)."""
        res = self.copy()
        res.qconj = -1
        res._set_charges(self.chinfo.make_valid(-self.charges))
        return res

    def sort(self, *args, **kwargs):
        return self.to_LegCharge().sort(*args, **kwargs)

    def bunch(self, *args, **kwargs):
        return self.to_LegCharge().bunch(*args, **kwargs)

    def project(self, *args, **kwargs):
        """Projecting a pipe yields a plain LegCharge (the pipe structure is lost)."""
        return self.to_LegCharge().project(*args, **kwargs)

    def map_incoming_flat(self, incoming_indices):
        """Map flat indices of the incoming legs to a flat index of the pipe (reference :1730)."""
        if len(incoming_indices) != self.nlegs:
            raise ValueError('wrong len of incoming_indices')
        qind_in = np.empty((1, self.nlegs), dtype=np.intp)
        within = np.empty(self.nlegs, dtype=np.intp)
        for li, (leg, idx) in enumerate(zip(self.legs, incoming_indices)):
            qi, w = leg.get_qindex(idx)
            qind_in[0, li] = qi
            within[li] = w
        row = self.q_map[self._map_incoming_qind(qind_in)[0]]
        sizes = [l.get_block_sizes()[qi] for l, qi in zip(self.legs, qind_in[0])]
        inner = 0
        for w, s in zip(within, sizes):
            inner = inner * int(s) + int(w)
        return int(self.slices[row[2]] + row[0] + inner)

    def _init_from_legs(self, sort=True, bunch=True):
        """Build charges, slices, q_map, q_map_slices (reference charges.py:1780 / pyx:545)."""
        nlegs = self.nlegs
        chinfo = self.chinfo
        qnumber = chinfo.qnumber
        subq = self.subqshape
        nblocks = int(np.prod(subq))
        # strides for row-major enumeration of the incoming qindex tuples
        strides = np.ones(nlegs, dtype=np.intp)
        for i in range(nlegs - 2, -1, -1):
            strides[i] = strides[i + 1] * subq[i + 1]
        self._strides = strides
        flat = np.arange(nblocks, dtype=np.intp)
        grid = np.empty((nblocks, nlegs), dtype=np.intp)
        for i in range(nlegs):
            grid[:, i] = (flat // strides[i]) % subq[i]
        q_map = np.empty((nblocks, 3 + nlegs), dtype=np.intp)
        q_map[:, 3:] = grid
        blocksizes = np.ones(nblocks, dtype=np.intp)
        charges = np.zeros((nblocks, qnumber), dtype=QTYPE)
        for i, leg in enumerate(self.legs):
            blocksizes *= leg.get_block_sizes()[grid[:, i]]
            if qnumber:
                charges += (self.qconj * leg.qconj) * leg.charges[grid[:, i]]
        if qnumber:
            charges = chinfo.make_valid(charges)
        if sort and qnumber > 0 and nblocks > 1:
            perm = _lexsort_rows(charges)
            q_map = q_map[perm]
            charges = charges[perm]
            blocksizes = blocksizes[perm]
            self._perm = _inverse_permutation(perm)
        else:
            self._perm = None
        self._set_charges(charges)
        self.sorted = bool(sort or qnumber == 0)
        self._set_block_sizes(blocksizes)
        q_map[:, 0] = self.slices[:-1]
        q_map[:, 1] = self.slices[1:]
        if bunch:
            idx = _row_change_points(charges)
            self._set_charges(charges[idx[:-1]])
            self._set_slices(self.slices[idx])
            q_map_Qi = np.zeros(nblocks, dtype=np.intp)
            q_map_Qi[idx[1:-1]] = 1
            q_map_Qi = np.cumsum(q_map_Qi)
            self.bunched = True
        else:
            q_map_Qi = np.arange(nblocks, dtype=np.intp)
            idx = np.arange(nblocks + 1, dtype=np.intp)
            self.bunched = self.is_bunched()
        q_map[:, 2] = q_map_Qi
        q_map[:, :2] -= self.slices[q_map_Qi][:, np.newaxis]
        self.q_map = q_map
        self.q_map_slices = idx

    def _map_incoming_qind(self, qind_incoming):
        """Rows of `q_map` belonging to the given incoming qindex tuples (reference :1860)."""
        inds = np.dot(np.asarray(qind_incoming, dtype=np.intp), self._strides)
        if self._perm is None:
            return inds
        return self._perm[inds]

    def __str__(self):
        return 'LegPipe(shape {0!s}->{1:d}, qconj {2}->{3:+d}; block numbers {4!s}->{5:d})'.format(
            self.subshape, self.ind_len, '(' + ', '.join('%+d' % l.qconj for l in self.legs) + ')',
            self.qconj, self.subqshape, self.block_number)

    def __repr__(self):
        return 'LegPipe({0!r},\nqconj={1:+d}, sort={2!r}, bunch={3!r})'.format(
            list(self.legs), self.qconj, self.sorted, self.bunched)
