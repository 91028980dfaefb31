"""dev check (NOT collected by pytest): the opt-in kernels prepared in round 1 without GPU time, switched on one by one
on the benchmark workload (short chain), with energies compared against the default configuration.

    python tests/dev_optins_gpu_check.py [chi=1024] [L=24]

configurations: default | mpo_apply='fused' (b200_mid_contract_f64) | + lanczos device_scalars
(b200_lanczos_update_dev_f64 / b200_scal_rsqrt_dev_f64) | + jacobi_eig_kernel_v2 (b200_svd_set_eig_variant(2)) |
identity_env (skip the identity components LP[IdL], RP[IdR]: D-1 instead of D large GEMMs per side; host logic only).
Prints one JSON line per configuration: seconds per sweep (CUDA events, 2nd of two sweeps), E, kernel family times.
"""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
from tenpy_b200 import backend


def main():
    chi = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    L = int(sys.argv[2]) if len(sys.argv) > 2 else 24
    lib = backend.get_lib()
    import bench
    from tenpy_b200.models import TFIChain
    from tenpy_b200.algorithms import dmrg
    model = TFIChain({'L': L, 'J': 1., 'g': 1., 'conserve': None})
    base = {'mixer': None, 'combine': True, 'diag_method': 'lanczos', 'svd_warm_start': False,
            'trunc_params': {'chi_max': chi, 'svd_min': 1e-45, 'trunc_cut': None}}
    configs = [('default', {}, 1),
               ('fused_mpo_apply', {'mpo_apply': 'fused'}, 1),
               ('fused+device_scalars', {'mpo_apply': 'fused', 'device_scalars': True}, 1),
               ('fused+device_scalars+eig_v2', {'mpo_apply': 'fused', 'device_scalars': True}, 2),
               ('identity_env', {'identity_env': True}, 1),
               ('identity_env+fused', {'identity_env': True, 'mpo_apply': 'fused'}, 1),
               ('identity_env+fused+device_scalars+eig_v2', {'identity_env': True, 'mpo_apply': 'fused',
                                                             'device_scalars': True}, 2)]
    E_ref = None
    for name, extra, eig_variant in configs:
        opts = dict(base)
        opts['lanczos_params'] = {'N_min': 10, 'N_max': 10, 'device_scalars': bool(extra.get('device_scalars', False))}
        if 'mpo_apply' in extra:
            opts['mpo_apply'] = extra['mpo_apply']
        if 'identity_env' in extra:
            opts['identity_env'] = True
        psi = bench.synthetic_mps(model, L, chi, 2, seed=0)
        eng = dmrg.TwoSiteDMRGEngine(psi, model, opts)
        old = lib.svd_set_eig_variant(eig_variant)
        try:
            eng.sweep()
            eng.sweep()
            torch.cuda.synchronize()
            lib.profile = {}
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            eng.sweep()
            ev1.record()
            torch.cuda.synchronize()
            fam = {k: round(v[1], 2) for k, v in lib.profile_summary().items()}
            lib.profile = None
        finally:
            lib.svd_set_eig_variant(old)
        E = float(eng.update_stats['E_total'][-1])
        if E_ref is None:
            E_ref = E
        print(json.dumps({'config': name, 'sweep_s': ev0.elapsed_time(ev1) / 1e3, 'E': E, 'dE_vs_default': E - E_ref,
                          'family_ms': fam}))
        assert abs(E - E_ref) < 1e-9 * abs(E_ref), 'energy differs from the default configuration'
    print('ok')


if __name__ == '__main__':
    main()
