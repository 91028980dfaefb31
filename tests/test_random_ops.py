"""Randomised differential test of the Array operations of the path against dense NumPy -- the reference's own unit-test
strategy (tests/test_np_conserved.py compares every operation with the result on `to_ndarray()`):
random charge rules (none, U(1), U(1)xZ2, Z3), random legs (sorted or not, repeated charges = unbunched sectors, ragged
block sizes), random total charges incl. ones that allow no block at all (empty Arrays).  Runs on the numpy test double
(host logic) and, marked ``gpu``, on the CUDA kernels through the C ABI."""
import numpy as np
import pytest


def _rand_leg(rng, npc, chinfo, n, qconj):
    q = chinfo.make_valid(rng.integers(-2, 3, size=(n, chinfo.qnumber)))
    if chinfo.qnumber and rng.random() < 0.6:
        q = q[np.lexsort(q.T)]
    return npc.LegCharge.from_qflat(chinfo, q, qconj)


def _rand_array(rng, npc, legs, qtotal=None, labels=None):
    return npc.Array.from_func(rng.standard_normal, legs, qtotal=qtotal, labels=labels)


def _check_ops(n_cases=40, seed=20260923):
    from tenpy_b200.linalg import np_conserved as npc
    rng = np.random.default_rng(seed)
    chinfos = [npc.ChargeInfo(), npc.ChargeInfo([1], ['N']), npc.ChargeInfo([1, 2], ['N', 'P']), npc.ChargeInfo([3], ['Z3'])]
    for case in range(n_cases):
        ci = chinfos[case % len(chinfos)]
        dims = [int(x) for x in rng.integers(1, 7, size=4)]
        legs = [_rand_leg(rng, npc, ci, d, int(rng.choice([-1, 1]))) for d in dims]
        qtot = ci.make_valid(rng.integers(-1, 2, size=ci.qnumber)) if rng.random() < 0.5 else None
        a = _rand_array(rng, npc, legs, qtot, labels=['a', 'b', 'c', 'd'])
        a.test_sanity()
        A = a.to_ndarray()
        tol = 1e-13 * max(1., np.abs(A).max())
        # --- transpose
        perm = [int(x) for x in rng.permutation(4)]
        t = a.transpose(perm)
        t.test_sanity()
        assert np.array_equal(t.to_ndarray(), A.transpose(perm))
        # --- combine two random (non adjacent) legs, split back
        i, j = (int(x) for x in rng.choice(4, size=2, replace=False))
        la, lb = a.get_leg_labels()[i], a.get_leg_labels()[j]
        c = a.combine_legs([la, lb])
        c.test_sanity()
        assert c.rank == 3 and np.isclose(npc.norm(c), np.linalg.norm(A))
        back = c.split_legs().transpose(['a', 'b', 'c', 'd'])
        assert np.array_equal(back.to_ndarray(), A)
        # dense check of the combined array through the pipe's index map
        pipe = c.get_leg('(%s.%s)' % (la, lb))
        rest = [x for x in range(4) if x not in (i, j)]
        dense_c = A.transpose([i, j] + rest).reshape((dims[i] * dims[j],) + tuple(dims[r] for r in rest))
        pos = c.get_leg_index('(%s.%s)' % (la, lb))
        flat = np.array([pipe.map_incoming_flat([x, y]) for x in range(dims[i]) for y in range(dims[j])])
        got = np.moveaxis(c.to_ndarray(), pos, 0)
        assert np.array_equal(got[flat], dense_c)
        # --- tensordot with a partner sharing two legs (conjugated), random extra legs
        k1, k2 = (int(x) for x in rng.choice(4, size=2, replace=False))
        extra = [_rand_leg(rng, npc, ci, int(rng.integers(1, 5)), int(rng.choice([-1, 1]))) for _ in range(2)]
        b_legs = [extra[0], legs[k1].conj(), extra[1], legs[k2].conj()]
        b = _rand_array(rng, npc, b_legs, ci.make_valid(rng.integers(-1, 2, size=ci.qnumber)), labels=['e', 'k1', 'f', 'k2'])
        r = npc.tensordot(a, b, axes=[[a.get_leg_labels()[k1], a.get_leg_labels()[k2]], ['k1', 'k2']])
        ref = np.tensordot(A, b.to_ndarray(), axes=[[k1, k2], [1, 3]])
        if isinstance(r, npc.Array):
            r.test_sanity()
            assert np.max(np.abs(r.to_ndarray() - ref)) <= 1e-12 * max(1., np.abs(ref).max()) * 8
        else:
            assert abs(r - ref) < 1e-12
        # --- inner / norm / linear combinations
        a2 = _rand_array(rng, npc, legs, a.qtotal, labels=['a', 'b', 'c', 'd'])
        assert abs(npc.inner(a, a2, axes='range', do_conj=True) - np.sum(A * a2.to_ndarray())) < 1e-12 * max(1., A.size)
        assert abs(npc.norm(a) - np.linalg.norm(A)) < 1e-12 * max(1., np.linalg.norm(A))
        lin = a + a2 * 0.5
        assert np.max(np.abs(lin.to_ndarray() - (A + 0.5 * a2.to_ndarray()))) <= tol * 4
        a3 = a.copy(deep=True).iadd_prefactor_other(-2., a2)
        assert np.max(np.abs(a3.to_ndarray() - (A - 2. * a2.to_ndarray()))) <= tol * 4
        # --- scale_axis / iproject / take_slice / add_leg
        ax = int(rng.integers(0, 4))
        s = rng.standard_normal(dims[ax])
        sc = a.scale_axis(s, ax)
        shape = [1] * 4
        shape[ax] = dims[ax]
        assert np.max(np.abs(sc.to_ndarray() - A * s.reshape(shape))) <= tol * 4
        mask = rng.random(dims[ax]) < 0.6
        if not np.any(mask):
            mask[int(rng.integers(0, dims[ax]))] = True
        pr = a.copy(deep=True)
        pr.iproject(mask, ax)
        pr.test_sanity()
        assert np.array_equal(pr.to_ndarray(), np.compress(mask, A, axis=ax))
        idx = int(rng.integers(0, dims[ax]))
        ts = a.take_slice(idx, ax)
        ts.test_sanity()
        assert np.array_equal(ts.to_ndarray(), np.take(A, idx, axis=ax))
        new_leg = _rand_leg(rng, npc, ci, 3, int(rng.choice([-1, 1])))
        jn = int(rng.integers(0, 3))
        al = a.add_leg(new_leg, jn, axis=ax, label='n')
        al.test_sanity()
        dense = np.zeros(A.shape[:ax] + (3,) + A.shape[ax:])
        sl = [slice(None)] * 5
        sl[ax] = jn
        dense[tuple(sl)] = A
        assert np.array_equal(al.to_ndarray(), dense)
        # --- concatenate along a random axis with a second array (new random leg there)
        legs2 = list(legs)
        legs2[ax] = _rand_leg(rng, npc, ci, int(rng.integers(1, 5)), legs[ax].qconj)
        a4 = _rand_array(rng, npc, legs2, a.qtotal, labels=['a', 'b', 'c', 'd'])
        cat = npc.concatenate([a, a4], axis=ax)
        cat.test_sanity()
        assert np.array_equal(cat.to_ndarray(), np.concatenate([A, a4.to_ndarray()], axis=ax))
        # --- svd / eigh / qr on the matrix (a b) x (c d): valid factorisations
        m = a.combine_legs([['a', 'b'], ['c', 'd']], qconj=[+1, -1])
        if m.stored_blocks:
            U, S, VH = npc.svd(m, inner_labels=['x', 'y'])
            rec = npc.tensordot(U.scale_axis(S, 1), VH, axes=1)
            assert npc.norm(rec - m) < 1e-12 * max(1., npc.norm(m))
            Sd = np.linalg.svd(m.to_ndarray(), compute_uv=False)
            Sd = np.sort(Sd[Sd > 1e-13])[::-1]
            Sg = np.sort(S[S > 1e-13])[::-1]
            assert len(Sd) == len(Sg) and np.max(np.abs(Sd - Sg)) < 1e-12 * max(1., Sd[0])
            Q, R = npc.qr(m, inner_labels=['x', 'y'])
            assert npc.norm(npc.tensordot(Q, R, axes=1) - m) < 1e-12 * max(1., npc.norm(m))
            qd = Q.to_ndarray()
            assert np.max(np.abs(qd.T @ qd - np.eye(qd.shape[1]))) < 1e-12
            rho = npc.tensordot(m, m.conj(), axes=[1, 1])
            W, V = npc.eigh(rho)
            vd = V.to_ndarray()
            assert np.max(np.abs(rho.to_ndarray() @ vd - vd * W)) < 1e-11 * max(1., np.abs(W).max())


def test_random_ops_host_logic(fake_device):
    _check_ops()


@pytest.mark.gpu
def test_random_ops_gpu(gpu_lib):
    _check_ops()
