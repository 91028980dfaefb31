#!/bin/bash
# round 2, call D: why is the int8 kernel at 25 % of the pipe?  layout rate probes, pipeline-depth A/B, ncu --set full
T=gpurun_out
mkdir -p $T
timeout 120 ./profiles/tc_i8_probe rate > $T/r02d_tc_rate.log 2>&1; cat $T/r02d_tc_rate.log
timeout 200 python profiles/ozaki_bench.py 1024 > $T/r02d_ozaki_cps4.jsonl 2> $T/r02d_ozaki_cps4.err
B200_OZ_CPS=2 timeout 200 python profiles/ozaki_bench.py 1024 > $T/r02d_ozaki_cps2.jsonl 2> $T/r02d_ozaki_cps2.err
python - <<'PY'
import json
for tag in ('cps4', 'cps2'):
    for line in open('gpurun_out/r02d_ozaki_%s.jsonl' % tag):
        d = json.loads(line)
        print(tag, d['shape'], 'dmma %.3f ms' % d['dmma_ms'], ' '.join('%s mm %.3f ms (%.0f Tops, %.1f TF)' % (k, d[k]['mm_ms'], d[k]['int8_Tops'], d[k]['mm_fp64_equiv_tflops']) for k in ('s7', 's8', 's9')))
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:oz_gemm -c 2 -o $T/r02d_oz_gemm python profiles/ozaki_one.py 7 2 > $T/r02d_ncu.log 2>&1; tail -3 $T/r02d_ncu.log
timeout 400 python -m pytest tests -m gpu -x -q > $T/r02d_gpu_tests.log 2>&1; tail -n 5 $T/r02d_gpu_tests.log
timeout 200 python bench.py --steps 1 --warmup 3 --no-cpu > $T/r02d_bench.json 2> $T/r02d_bench.err; tail -c 300 $T/r02d_bench.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02d_bench.json').read().strip().splitlines()[-1])
print('bench sweep_s', d['value'], 'e2e', d['e2e']['value'], 'E', d['result']['E'], d['kernel_family_ms_per_sweep'])
PY
