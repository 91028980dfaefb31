#!/bin/bash
# round 2, call G: dense matvec recipe replay in the bench, tcgen05.ld 16x256b fragment layout, inner sweeps of the pivot
# solver, configs[2]/[3] parity at L=64/32 chi=256 against the reference goldens
T=gpurun_out
mkdir -p $T
timeout 60 ./profiles/tc_i8_probe ldshape > $T/r02g_ldshape.log 2>&1; cat $T/r02g_ldshape.log
timeout 300 python profiles/svd_variants.py > $T/r02g_svd_variants.jsonl 2> $T/r02g_svd_variants.err; cat $T/r02g_svd_variants.jsonl; tail -c 300 $T/r02g_svd_variants.err
timeout 900 python -m pytest tests/test_large_parity.py tests/test_tebd.py tests/test_qr_truncation.py tests/test_gpu_kernels.py -m gpu -x -q > $T/r02g_tests.log 2>&1; tail -n 12 $T/r02g_tests.log
timeout 400 python bench.py --steps 1 --warmup 3 > $T/r02g_bench.json 2> $T/r02g_bench.err; tail -c 400 $T/r02g_bench.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02g_bench.json').read().strip().splitlines()[-1])
print('bench sweep_s', d['value'], 'e2e', d['e2e']['value'], 'E', d['result']['E'], d['kernel_family_ms_per_sweep'], 'launches', d['gpu_launches'])
print('parity', d['parity']); print('default lanczos', d['reference_default_lanczos']); print('cpu', {k: d['cpu_baseline'][k] for k in ('value', 'kind', 'cores', 'per_bond_s', 'matvec_s', 'svd_s') if k in d['cpu_baseline']})
PY
