"""CPU checks of device code written as host-callable per-thread functions.

(1) the version-2 pivot eigen-solver of the Jacobi SVD (tenpy_b200/csrc/jacobi_eig_core.cuh): the phase
functions the CUDA kernel `jacobi_eig_kernel_v2` calls between barriers are compiled for the host and run thread by
thread (tests/csrc/eig_core_host.cpp), next to a sequential restatement of the GPU-verified version 1.
(2) the per-thread body of `mid_contract_kernel` (tenpy_b200/csrc/mid_contract_core.cuh) against a triple loop.
(3) the phases of the Householder `block_qr_kernel` (tenpy_b200/csrc/block_qr_core.cuh)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which('g++') is None, reason='needs g++')
def test_eig_core_phases_on_host(tmp_path):
    exe = str(tmp_path / 'eig_core_host')
    src = os.path.join(ROOT, 'tests', 'csrc', 'eig_core_host.cpp')
    subprocess.check_call(['g++', '-O2', '-std=c++17', '-Wall', '-o', exe, src])
    out = subprocess.run([exe, '60'], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr
    assert out.stdout.strip().endswith('ok')


@pytest.mark.skipif(shutil.which('g++') is None, reason='needs g++')
def test_mid_contract_column_on_host(tmp_path):
    exe = str(tmp_path / 'mid_contract_host')
    src = os.path.join(ROOT, 'tests', 'csrc', 'mid_contract_host.cpp')
    subprocess.check_call(['g++', '-O2', '-std=c++17', '-Wall', '-o', exe, src])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr
    assert out.stdout.strip().endswith('ok')


@pytest.mark.skipif(shutil.which('g++') is None, reason='needs g++')
def test_block_qr_phases_on_host(tmp_path):
    exe = str(tmp_path / 'block_qr_host')
    src = os.path.join(ROOT, 'tests', 'csrc', 'block_qr_host.cpp')
    subprocess.check_call(['g++', '-O2', '-std=c++17', '-Wall', '-o', exe, src])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr
    assert out.stdout.strip().endswith('ok')
