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
table = scan.run_scan(configs, run, cost_fn=lambda c: c['chi'] ** 3, schedule=os.environ.get('SCAN_SCHEDULE', 'dynamic'))
if dist.get_rank() == 0:
    np.save({out!r}, table)
dist.destroy_process_group()
'''

SCHED_WORKER = r'''
import os, sys, time
sys.path.insert(0, {root!r})
import numpy as np, torch, torch.distributed as dist
dist.init_process_group('gloo')
from tenpy_b200 import scan
rank = dist.get_rank()
# run i "takes" dur[i] seconds; the cost estimate (what a static assignment sees) is the same for the two big ones
dur = [0.9, 0.3, 0.3, 0.3, 0.1, 0.1]
est = [8., 8., 1., 1., 1., 1.]
def run(cfg):
    if cfg['i'] == int(os.environ.get('FAIL_RUN', '-1')):
        raise ValueError('boom')
    time.sleep(dur[cfg['i']])
    return [dur[cfg['i']], float(rank)]
t0 = time.time()
try:
    table = scan.run_scan([dict(i=i) for i in range(6)], run, cost_fn=lambda c: est[c['i']],
                          schedule=os.environ.get('SCAN_SCHEDULE', 'dynamic'))
    table2 = scan.run_scan([dict(i=i) for i in range(6)], run, cost_fn=lambda c: est[c['i']])      # a second scan: fresh counter
    if rank == 0:
        np.save({out!r}, np.concatenate([table, table2, [[time.time() - t0, 0., 0.]]]))
except RuntimeError as e:
    open({out!r} + '.err%d' % rank, 'w').write(str(e))
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


def _launch(tmp_path, script_text, port, env_extra=None, nproc=2):
    out = str(tmp_path / 'table.npy')
    script = tmp_path / 'worker.py'
    script.write_text(script_text.format(root=ROOT, out=out))
    env = dict(os.environ, MASTER_ADDR='127.0.0.1', **(env_extra or {}))
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=%d' % nproc, '--master-addr',
           '127.0.0.1', '--master-port', str(port), str(script)]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    return out


def test_scan_dynamic_schedule_balances_on_measured_times(tmp_path):
    """pull-based scheduling over the rendezvous store: every run exactly once, ranks balanced by the times the runs really
    take (the static LPT assignment on the estimates puts 0.9 + 3 x ... on one rank), a second scan gets a fresh counter"""
    out = _launch(tmp_path, SCHED_WORKER, 29619)
    res = np.load(out)
    for table in (res[:6], res[6:12]):
        assert np.array_equal(table[:, 0], np.arange(6.))
        busy = [table[table[:, 2] == r, 1].sum() for r in (0., 1.)]
        assert abs(sum(busy) - 2.0) < 1e-9 and max(busy) <= 1.11, busy       # optimum 1.0; static LPT on `est`: 1.3 / 0.7
    out = _launch(tmp_path, SCHED_WORKER, 29621, {'SCAN_SCHEDULE': 'static'})
    table = np.load(out)[:6]
    busy = [table[table[:, 2] == r, 1].sum() for r in (0., 1.)]
    assert abs(max(busy) - 1.3) < 1e-9, busy


def test_scan_failing_run_raises_on_every_rank(tmp_path):
    """a run that raises does not leave the other ranks waiting in a collective: all of them get a RuntimeError"""
    out = _launch(tmp_path, SCHED_WORKER, 29623, {'FAIL_RUN': '2'})
    assert not os.path.exists(out)
    msgs = [open(out + '.err%d' % r).read() for r in (0, 1)]
    assert all('1 run(s) failed' in m for m in msgs) and any('boom' in m for m in msgs)


def test_assign_runs():
    from tenpy_b200 import scan
    out = scan.assign_runs([8, 1, 1, 1, 1, 1, 1, 1, 1], 2)
    assert sorted(sum(out, [])) == list(range(9))
    assert out[0] == [0] or out[1] == [0]
