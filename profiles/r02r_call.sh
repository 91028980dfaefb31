#!/bin/bash
# round 2, call R: pivot eigen-solver modes (2 inner sweeps / 1 / cross) and the single-launch rounds in whole DMRG sweeps
T=gpurun_out
mkdir -p $T
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_large_parity.py -m gpu -q > $T/r02r_tests.log 2>&1; tail -n 3 $T/r02r_tests.log
timeout 600 python profiles/svd_variants.py > $T/r02r_svd_variants.jsonl 2> $T/r02r_svd_variants.err
python - <<'PY'
import json
for line in open('gpurun_out/r02r_svd_variants.jsonl'):
    d = json.loads(line)
    print(d['case'], {k: v for k, v in d.items() if k.endswith('_ms') or k.endswith('_sweeps')})
PY
timeout 300 python profiles/svd_gap_probe.py 40 1024 > $T/r02r_svd_gap.json 2> $T/r02r_svd_gap.err; cat $T/r02r_svd_gap.json | cut -c1-900; tail -c 300 $T/r02r_svd_gap.err
for inner in 2 0 1; do
  B200_SVD_INNER=$inner timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --no-blocksparse > $T/r02r_tfi_in$inner.json 2> $T/r02r_tfi_in$inner.err
  python -c "
import json; d=json.load(open('$T/r02r_tfi_in$inner.json')); print('tfi inner', $inner, d['value'], d['kernel_family_ms_per_sweep'], d['result']['svd_jacobi_sweeps_mean'], d['parity']['E_rel_err'])"
done
for cfg in "2 512" "0 512" "2 0" "0 0"; do
  set -- $cfg
  B200_SVD_INNER=$1 B200_SVD_FUSED_LD=$2 timeout 600 python bench.py --workload xxz --steps 2 --warmup 1 > $T/r02r_xxz_in$1_ld$2.json 2> $T/r02r_xxz_in$1_ld$2.err
  python -c "
import json; d=json.load(open('$T/r02r_xxz_in$1_ld$2.json')); print('xxz inner', $1, 'fused ld', $2, d['value'], d.get('kernel_family_ms_per_sweep'), d['result'])"
done
for inner in 2 0; do
  B200_SVD_INNER=$inner timeout 900 python bench.py --workload hubbard --steps 1 --warmup 1 > $T/r02r_hubbard_in$inner.json 2> $T/r02r_hubbard_in$inner.err
  python -c "
import json; d=json.load(open('$T/r02r_hubbard_in$inner.json')); print('hubbard inner', $inner, d['value'], d.get('kernel_family_ms_per_sweep'), d['result'])"
done
