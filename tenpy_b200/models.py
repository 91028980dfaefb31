"""The three benchmark chains as explicit MPOs (cold path, host side).

The reference builds these through its generic `CouplingMPOModel` machinery (``tenpy/models/model.py``,
``tf_ising.py:74`` `TFIChain`, ``spins.py`` `SpinChain`, ``hubbard.py:207`` `FermiHubbardChain`); model
construction is outside the hot path, so here each Hamiltonian is written
down directly as its finite-state-machine MPO.  The resulting bond dimensions equal the reference's
(TFI D=3, XXZ D=5, Hubbard D=6) and energies are pinned against the reference in tests/golden.

* TFIChain:          ``H = -J sum_i sx_i sx_{i+1} - g sum_i sz_i``  (Pauli matrices)
* SpinChain (S=1/2): ``H = sum_i Jx Sx Sx + Jy Sy Sy + Jz Sz Sz - hz sum_i Sz_i``
* FermiHubbardChain: ``H = -t sum_{i,s} (c^+_{i,s} c_{i+1,s} + h.c.) + U sum_i n_up n_dn - mu sum_i n_i``
"""
# Copyright (C) 2026 tenpy_b200 authors. Apache-2.0.

import numpy as np

from .networks.site import SpinHalfSite, SpinHalfFermionSite
from .networks.mpo import MPO

__all__ = ['TFIChain', 'SpinChain', 'XXZChain', 'FermiHubbardChain']


class _ChainModel:
    """holds `lat_sites` and `H_MPO` (the attributes the DMRG engine reads from a reference model)"""

    def __init__(self, sites, grid, op_names_first_row, IdL, IdR):
        self.sites = sites
        self.L = len(sites)
        site = sites[0]
        chinfo = site.leg.chinfo
        D = len(grid)
        if chinfo.qnumber:
            # charge of MPO index b = charge of the operator in W[0, b] (IdL row); see mpo.MPO.from_grids
            wq = np.zeros((D, chinfo.qnumber), dtype=np.int64)
            for b, name in enumerate(op_names_first_row):
                if name is not None:
                    wq[b] = site.op_charge(name)
            w_charges = [chinfo.make_valid(wq)] * (self.L + 1)
        else:
            w_charges = [None] * (self.L + 1)
        self.H_MPO = MPO.from_grids(sites, [grid] * self.L, w_charges, IdL, IdR)
        self.lat_sites = sites


class TFIChain(_ChainModel):
    """Transverse field Ising chain (reference tf_ising.py:74); params ``L, J, g, conserve``."""

    def __init__(self, params):
        L = params['L']
        J, g = params.get('J', 1.), params.get('g', 1.)
        conserve = params.get('conserve', None)
        if params.get('bc_MPS', 'finite') != 'finite':
            raise NotImplementedError('finite chains only')
        site = SpinHalfSite(conserve=conserve)
        grid = [[[(1., 'Id')], [(1., 'Sigmax')], [(-g, 'Sigmaz')]],
                [None, None, [(-J, 'Sigmax')]],
                [None, None, [(1., 'Id')]]]
        _ChainModel.__init__(self, [site] * L, grid, ['Id', 'Sigmax', None], 0, 2)


class SpinChain(_ChainModel):
    """Spin-1/2 chain (reference spins.py `SpinChain`, S=0.5); params ``L, Jx, Jy, Jz, hz, conserve``."""

    def __init__(self, params):
        L = params['L']
        if params.get('S', 0.5) != 0.5:
            raise NotImplementedError('S=1/2 only')
        Jx, Jy, Jz = params.get('Jx', 1.), params.get('Jy', 1.), params.get('Jz', 1.)
        hz = params.get('hz', 0.)
        conserve = params.get('conserve', 'best')
        if conserve == 'best':
            conserve = 'Sz' if Jx == Jy else 'parity'
        if conserve == 'Sz' and Jx != Jy:
            raise ValueError('Sz is not conserved for Jx != Jy')
        site = SpinHalfSite(conserve=conserve)
        jpm = (Jx + Jy) / 4.
        jpp = (Jx - Jy) / 4.
        last = [[(-hz, 'Sz')] if hz != 0. else None,
                [(jpm, 'Sm')] + ([(jpp, 'Sp')] if jpp != 0. else []),
                [(jpm, 'Sp')] + ([(jpp, 'Sm')] if jpp != 0. else []),
                [(Jz, 'Sz')],
                [(1., 'Id')]]
        grid = [[[(1., 'Id')], [(1., 'Sp')], [(1., 'Sm')], [(1., 'Sz')], last[0]],
                [None, None, None, None, last[1]],
                [None, None, None, None, last[2]],
                [None, None, None, None, last[3]],
                [None, None, None, None, last[4]]]
        _ChainModel.__init__(self, [site] * L, grid, ['Id', 'Sp', 'Sm', 'Sz', None], 0, 4)


class XXZChain(SpinChain):
    """``H = Jxx/2 (S+S- + h.c.) + Jz Sz Sz - hz Sz`` (reference xxz_chain.py)."""

    def __init__(self, params):
        p = dict(params)
        Jxx = p.pop('Jxx', 1.)
        p.setdefault('Jz', 1.)
        p['Jx'] = p['Jy'] = Jxx
        p.setdefault('conserve', 'Sz')
        SpinChain.__init__(self, p)


class FermiHubbardChain(_ChainModel):
    """Spinful Fermi-Hubbard chain (reference hubbard.py:207); params ``L, t, U, mu, cons_N, cons_Sz``.

    Jordan-Wigner: ``c^+_i c_{i+1} = (Cd JW)_i (C)_{i+1}``, ``c^+_{i+1} c_i = (JW C)_i (Cd)_{i+1}``."""

    def __init__(self, params):
        L = params['L']
        t, U, mu = params.get('t', 1.), params.get('U', 0.), params.get('mu', 0.)
        site = SpinHalfFermionSite(cons_N=params.get('cons_N', 'N'), cons_Sz=params.get('cons_Sz', 'Sz'))
        d = site.get_dense
        site.add_op('CduJW', d('Cdu') @ d('JW'), permute=False)
        site.add_op('JWCu', d('JW') @ d('Cu'), permute=False)
        site.add_op('CddJW', d('Cdd') @ d('JW'), permute=False)
        site.add_op('JWCd', d('JW') @ d('Cd'), permute=False)
        onsite = [(U, 'NuNd')] + ([(-mu, 'Ntot')] if mu != 0. else [])
        if U == 0. and mu == 0.:
            onsite = None
        grid = [[[(1., 'Id')], [(1., 'CduJW')], [(1., 'JWCu')], [(1., 'CddJW')], [(1., 'JWCd')], onsite],
                [None, None, None, None, None, [(-t, 'Cu')]],
                [None, None, None, None, None, [(-t, 'Cdu')]],
                [None, None, None, None, None, [(-t, 'Cd')]],
                [None, None, None, None, None, [(-t, 'Cdd')]],
                [None, None, None, None, None, [(1., 'Id')]]]
        _ChainModel.__init__(self, [site] * L, grid, ['Id', 'CduJW', 'JWCu', 'CddJW', 'JWCd', None], 0, 5)
