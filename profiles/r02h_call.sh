#!/bin/bash
# round 2, call H: fragment-layout epilogue, pivot solver v3 / 1 inner sweep by default, deflation tolerance <= svd_min,
# configs[2]/[3] parity, the sweep driven by the reference's engine class
T=gpurun_out
mkdir -p $T
B200_OZ_DEBUG=1 timeout 120 python profiles/ozaki_one.py 7 2 > $T/r02h_oz_debug.log 2>&1; cat $T/r02h_oz_debug.log
timeout 200 python profiles/ozaki_bench.py 1024 > $T/r02h_ozaki.jsonl 2> $T/r02h_ozaki.err
python - <<'PY'
import json
for line in open('gpurun_out/r02h_ozaki.jsonl'):
    d = json.loads(line)
    print(d['shape'], 'dmma %.3f ms' % d['dmma_ms'], ' '.join('%s mm %.3f ms (%.0f Tops, %.1f TF)' % (k, d[k]['mm_ms'], d[k]['int8_Tops'], d[k]['mm_fp64_equiv_tflops']) for k in ('s7', 's8', 's9')))
PY
timeout 900 python -m pytest tests/test_ozaki.py tests/test_large_parity.py tests/test_gpu_kernels.py tests/test_gpu_parity.py -m gpu -x -q > $T/r02h_tests.log 2>&1; tail -n 12 $T/r02h_tests.log
timeout 500 python bench.py --steps 1 --warmup 3 > $T/r02h_bench.json 2> $T/r02h_bench.err; tail -c 400 $T/r02h_bench.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02h_bench.json').read().strip().splitlines()[-1])
print('bench sweep_s', d['value'], 'e2e', d['e2e']['value'], 'E', d['result']['E'], d['kernel_family_ms_per_sweep'], 'launches', d['gpu_launches'])
print('ab', d['ab'], 'default lanczos', d['reference_default_lanczos']); print('reference_driver', d.get('reference_driver'))
print('cpu', {k: d['cpu_baseline'][k] for k in ('value', 'kind', 'cores', 'per_bond_s', 'matvec_s', 'svd_s') if k in d['cpu_baseline']})
PY
