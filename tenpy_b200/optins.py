"""Switches for the kernels and contraction routes that are compiled and host-checked but have not run on a GPU yet
("opt-in kernels").  Off by default; the environment variable ``B200_OPTINS`` turns them on for a
whole process without code changes, so that the first GPU call of a round can A/B them:

    B200_OPTINS=all python bench.py ...            # everything
    B200_OPTINS=fused,identity python bench.py ...  # a subset

=========== =========================================================================================
name        effect
=========== =========================================================================================
fused       TwoSiteH.mpo_apply = 'fused'            (b200_mid_contract_f64 / b200_mid_contract2_f64)
identity    TwoSiteH.identity_env = True             (default since the end of round 1; B200_OPTINS name kept)
devscal     Lanczos option device_scalars = True     (b200_lanczos_update_dev_f64, b200_scal_rsqrt_dev_f64)
eigv2       b200_svd_set_eig_variant(2)              (jacobi_eig_kernel_v2)
qrhh        np_conserved.qr_method = 'householder'   (b200_block_qr_f64)
=========== =========================================================================================
"""
import os

KNOWN = ('fused', 'identity', 'devscal', 'eigv2', 'qrhh')


def requested(env=None):
    raw = (os.environ.get('B200_OPTINS', '') if env is None else env).strip().lower()
    if not raw:
        return ()
    names = KNOWN if raw == 'all' else tuple(x.strip() for x in raw.split(',') if x.strip())
    unknown = [n for n in names if n not in KNOWN]
    if unknown:
        raise ValueError('B200_OPTINS: unknown name(s) {0}; known: {1}'.format(unknown, KNOWN))
    return names


def apply(names=None, lib=None):
    """Apply the switches (default: those of ``B200_OPTINS``).  `lib`: the device library, for ``eigv2``.
    Returns the tuple of names applied."""
    names = requested() if names is None else tuple(names)
    if not names:
        return names
    from .algorithms import mps_common
    from .linalg import krylov_based, np_conserved
    if 'fused' in names:
        mps_common.TwoSiteH.mpo_apply = 'fused'
    if 'identity' in names:
        mps_common.TwoSiteH.identity_env = True
    if 'devscal' in names:
        krylov_based.DEVICE_SCALARS_DEFAULT = True
    if 'qrhh' in names:
        np_conserved.qr_method = 'householder'
    if 'eigv2' in names and lib is not None and hasattr(lib, 'svd_set_eig_variant'):
        lib.svd_set_eig_variant(2)
    return names
