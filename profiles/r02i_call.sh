#!/bin/bash
# round 2, call I: int8 kernel with top-aligned passes (transposition epilogue), per-bond MPO cache, configs[2]/[3] end to end
T=gpurun_out
mkdir -p $T
timeout 200 python profiles/ozaki_bench.py 1024 > $T/r02i_ozaki.jsonl 2> $T/r02i_ozaki.err
python - <<'PY'
import json
for line in open('gpurun_out/r02i_ozaki.jsonl'):
    d = json.loads(line)
    print(d['shape'], 'dmma %.3f ms' % d['dmma_ms'], ' '.join('%s mm %.3f ms (%.0f Tops, %.1f TF)' % (k, d[k]['mm_ms'], d[k]['int8_Tops'], d[k]['mm_fp64_equiv_tflops']) for k in ('s7', 's8', 's9')))
PY
timeout 600 python -m pytest tests/test_ozaki.py tests/test_gpu_kernels.py tests/test_gpu_parity.py -m gpu -x -q > $T/r02i_tests.log 2>&1; tail -n 4 $T/r02i_tests.log
timeout 400 python bench.py --steps 1 --warmup 3 --no-cpu > $T/r02i_bench.json 2> $T/r02i_bench.err; tail -c 300 $T/r02i_bench.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02i_bench.json').read().strip().splitlines()[-1])
print('bench sweep_s', d['value'], 'e2e', d['e2e']['value'], 'E', d['result']['E'], d['kernel_family_ms_per_sweep'], 'launches', d['gpu_launches'])
print('ab', d['ab'], 'default lanczos', d['reference_default_lanczos']); print('reference_driver', d.get('reference_driver', {}).get('value'))
PY
timeout 500 python bench.py --workload xxz --steps 2 --warmup 1 > $T/r02i_xxz.json 2> $T/r02i_xxz.err; tail -c 300 $T/r02i_xxz.err
timeout 700 python bench.py --workload hubbard --steps 1 --warmup 1 --ramp 7 > $T/r02i_hubbard.json 2> $T/r02i_hubbard.err; tail -c 300 $T/r02i_hubbard.err
python - <<'PY'
import json
for w in ('xxz', 'hubbard'):
    try:
        d = json.loads(open('gpurun_out/r02i_%s.json' % w).read().strip().splitlines()[-1])
        print(w, 'sweep_s', d['value'], 'ramp', d['ramp_sweep_s'], d['kernel_family_ms_per_sweep'], 'wall', d['host_wall_s_profiled_sweep'], d['structure'], d['result'])
        for h in d['gemm_by_flops']: print('   ', h)
    except Exception as e:
        print(w, 'no result', e)
PY
