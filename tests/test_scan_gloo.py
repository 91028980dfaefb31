"""CPU test of the multi-rank path (gloo, world size 2): sharding of independent DMRG runs over ranks,
broadcast of the model template, all-gather of the results.  Each rank runs a small DMRG through the numpy
test double of the device library (host-logic test; the GPU ranks use the real kernels)."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, 'tests'))
import numpy as np, torch, torch.distributed as dist
dist.init_process_group('gloo')
from tenpy_b200 import backend, scan
from fake_device import FakeDeviceLib
backend.use_library(FakeDeviceLib())
from tenpy_b200.models import TFIChain
from tenpy_b200.networks.mps import MPS
from tenpy_b200.algorithms import dmrg
tmpl = scan.broadcast_template([1.0, 8.0] if dist.get_rank() == 0 else [0.0, 0.0])
J, L = float(tmpl[0]), int(tmpl[1])
configs = [dict(g=0.5, chi=8), dict(g=1.5, chi=16), dict(g=1.0, chi=12)]
def run(cfg):
    M = TFIChain(dict(L=L, J=J, g=cfg['g'], conserve=None))
    psi = MPS.from_product_state(M.lat_sites, ['up'] * L)
    res = dmrg.run(psi, M, dict(mixer=None, max_E_err=1e-10, trunc_params=dict(chi_max=cfg['chi'], svd_min=1e-10)))
    return [res['E'], float(max(psi.chi)), float(dist.get_rank())]
table = scan.run_scan(configs, run, cost_fn=lambda c: c['chi'] ** 3)
if dist.get_rank() == 0:
    np.save({out!r}, table)
dist.destroy_process_group()
'''


def test_scan_two_ranks(tmp_path):
    out = str(tmp_path / 'table.npy')
    script = tmp_path / 'worker.py'
    script.write_text(WORKER.format(root=ROOT, out=out))
    env = dict(os.environ, MASTER_ADDR='127.0.0.1')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2', '--master-addr',
           '127.0.0.1', '--master-port', '29617', str(script)]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    table = np.load(out)
    assert table.shape == (3, 4)
    assert np.array_equal(table[:, 0], [0., 1., 2.])
    # exact ground state energies of the open TFI chain (free fermions), L=8
    from oracle import dmrg_dense as od
    for row, g in zip(table, [0.5, 1.5, 1.0]):
        ref = od.run_dmrg(od.tfi_mpo(g, 1.), 8, 2, [0] * 8, dict(chi_max=16, svd_min=1e-10), {}, max_E_err=1e-10)
        assert abs(row[1] - ref['E']) < 1e-8 * abs(ref['E'])
    # LPT assignment: the most expensive run (chi=16) alone on one rank
    ranks = table[:, 3]
    assert ranks[1] != ranks[0] or ranks[1] != ranks[2]
    assert set(ranks) == {0., 1.}


def test_assign_runs():
    from tenpy_b200 import scan
    out = scan.assign_runs([8, 1, 1, 1, 1, 1, 1, 1, 1], 2)
    assert sorted(sum(out, [])) == list(range(9))
    assert out[0] == [0] or out[1] == [0]
