#!/bin/bash
# round 2, call R: pivot eigen-solver modes (2 inner sweeps / 1 / cross) in whole DMRG sweeps
T=gpurun_out
mkdir -p $T
for inner in 2 0 1; do
  B200_SVD_INNER=$inner timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --no-blocksparse > $T/r02r_tfi_in$inner.json 2> $T/r02r_tfi_in$inner.err
  python -c "
import json; d=json.load(open('$T/r02r_tfi_in$inner.json')); print('tfi inner', $inner, d['value'], d['kernel_family_ms_per_sweep'], d['result']['svd_jacobi_sweeps_mean'], d['parity']['E_rel_err'])"
done
for inner in 2 0; do
  B200_SVD_INNER=$inner timeout 600 python bench.py --workload xxz --steps 2 --warmup 1 > $T/r02r_xxz_in$inner.json 2> $T/r02r_xxz_in$inner.err
  python -c "
import json; d=json.load(open('$T/r02r_xxz_in$inner.json')); print('xxz inner', $inner, d['value'], d.get('kernel_family_ms_per_sweep'), d['result'])"
done
for inner in 2 0; do
  B200_SVD_INNER=$inner timeout 900 python bench.py --workload hubbard --steps 1 --warmup 1 > $T/r02r_hubbard_in$inner.json 2> $T/r02r_hubbard_in$inner.err
  python -c "
import json; d=json.load(open('$T/r02r_hubbard_in$inner.json')); print('hubbard inner', $inner, d['value'], d.get('kernel_family_ms_per_sweep'), d['result'])"
done
timeout 600 python -m pytest tests/test_large_parity.py -m gpu -q > $T/r02r_tests.log 2>&1; tail -n 3 $T/r02r_tests.log
