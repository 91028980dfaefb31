#!/bin/bash
# round 2, call E: int8 kernel cycle split (B200_OZ_DEBUG), cheaper descriptor arithmetic; pivot eigen-solver v3
T=gpurun_out
mkdir -p $T
B200_OZ_DEBUG=1 timeout 120 python profiles/ozaki_one.py 7 2 > $T/r02e_oz_debug.log 2>&1; cat $T/r02e_oz_debug.log
B200_OZ_DEBUG=1 timeout 120 python profiles/ozaki_one.py 8 2 >> $T/r02e_oz_debug.log 2>&1; tail -3 $T/r02e_oz_debug.log
timeout 200 python profiles/ozaki_bench.py 1024 > $T/r02e_ozaki.jsonl 2> $T/r02e_ozaki.err
python - <<'PY'
import json
for line in open('gpurun_out/r02e_ozaki.jsonl'):
    d = json.loads(line)
    print(d['shape'], 'dmma %.3f ms' % d['dmma_ms'], ' '.join('%s mm %.3f ms (%.0f Tops, %.1f TF)' % (k, d[k]['mm_ms'], d[k]['int8_Tops'], d[k]['mm_fp64_equiv_tflops']) for k in ('s7', 's8', 's9')))
PY
timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_ozaki.py -m gpu -x -q > $T/r02e_tests.log 2>&1; tail -n 4 $T/r02e_tests.log
timeout 300 python profiles/svd_variants.py > $T/r02e_svd_variants.jsonl 2> $T/r02e_svd_variants.err; cat $T/r02e_svd_variants.jsonl; tail -c 300 $T/r02e_svd_variants.err
