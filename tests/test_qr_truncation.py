"""QR based truncation (reference truncation.py:370-713): `decompose_theta_qr_based` against the
reference's outputs (tests/golden/qr_trunc.npz from tests/golden/make_golden_qr_trunc.py) -- singular values,
truncation error, renormalisation and the reconstructed (gauge invariant) theta; isometry of the returned tensors."""
import numpy as np
import pytest

import helpers as h


def _check():
    from tenpy_b200.linalg import np_conserved as npc
    from tenpy_b200.linalg.charges import LegCharge
    from tenpy_b200.linalg.truncation import decompose_theta_qr_based
    g = h.load('qr_trunc.npz')
    for i in range(int(g['n'])):
        key = 'q%d' % i
        move_right, eig = bool(g[key + '_move_right']), bool(g[key + '_eig'])
        theta = h.to_product(h.oarray_from(g, key + '_theta'))
        old_leg = LegCharge.from_qind(theta.chinfo, g[key + '_oldleg_slices'], g[key + '_oldleg_charges'],
                                      int(g[key + '_oldleg_qconj']))
        tp = dict(chi_max=12, svd_min=1e-10)
        T_L, S, T_R, form, err, renorm = decompose_theta_qr_based(g[key + '_qL'], g[key + '_qR'], old_leg, theta,
                                                                  move_right, 0.5, 1, eig, tp, True, True)
        assert list(form) == [str(x) for x in g[key + '_form']]
        assert len(S) == len(g[key + '_S'])
        assert np.max(np.abs(np.sort(S) - np.sort(g[key + '_S']))) < (1e-7 if eig else 1e-10)   # eigh: sqrt of eigenvalues
        assert abs(renorm - g[key + '_renorm']) < 1e-10 * g[key + '_renorm']
        assert abs(err.eps - g[key + '_eps']) < 1e-12 + 1e-6 * g[key + '_eps']
        approx = npc.tensordot(T_L, T_R, axes=['vR', 'vL']) if eig else \
            npc.tensordot(T_L.scale_axis(S, 'vR'), T_R, axes=['vR', 'vL'])
        approx.ireplace_labels(['(vL.p)', '(p.vR)'], ['(vL.p0)', '(p1.vR)'])
        h.assert_close(h.to_oracle(approx), h.oarray_from(g, key + '_approx'), 1e-9, structure=False)
        if form[0] == 'A':
            iso = npc.tensordot(T_L.conj(), T_L, axes=['(vL*.p*)', '(vL.p)']).to_ndarray()
            assert np.max(np.abs(iso - np.eye(len(iso)))) < 1e-11
        if form[1] == 'B':
            iso = npc.tensordot(T_R, T_R.conj(), axes=['(p.vR)', '(p*.vR*)']).to_ndarray()
            assert np.max(np.abs(iso - np.eye(len(iso)))) < 1e-11
        # without error computation / second tensor
        T_L2, S2, T_R2, form2, err2, _ = decompose_theta_qr_based(g[key + '_qL'], g[key + '_qR'], old_leg, theta,
                                                                  move_right, 0.5, 1, eig, tp, False, False)
        assert np.isnan(err2.eps) and ((T_R2 is None) if move_right else (T_L2 is None))
        assert np.max(np.abs(np.sort(S2) - np.sort(S))) < 1e-13


def test_qr_based_truncation_host_logic(fake_device):
    _check()


@pytest.mark.gpu
def test_qr_based_truncation_gpu(gpu_lib):
    _check()
