#!/bin/bash
# round 2, call O: single-launch split kernel (parity + timing against the two-pass kernels), host-side profile of a sweep
T=gpurun_out
mkdir -p $T
timeout 600 python -m pytest tests/test_ozaki.py tests/test_large_parity.py tests/test_dropin_engine.py -m gpu -q > $T/r02o_tests.log 2>&1; tail -n 5 $T/r02o_tests.log
for v in fused twopass; do
  if [ $v = twopass ]; then export B200_OZ_SPLIT2=1; else unset B200_OZ_SPLIT2; fi
  timeout 200 python profiles/ozaki_bench.py 1024 > $T/r02o_ozaki_$v.jsonl 2> $T/r02o_ozaki_$v.err
  python - <<PY
import json
for line in open('gpurun_out/r02o_ozaki_$v.jsonl'):
    d = json.loads(line)
    print('$v', d['shape'], ' '.join('%s split A %.1f us B %.1f us mm %.3f ms err %.1e' % (k, d[k]['split_A_ms'] * 1e3, d[k]['split_B_ms'] * 1e3, d[k]['mm_ms'], d[k]['max_abs_diff_vs_dmma_rel']) for k in ('s7', 's8', 's9')))
PY
done
unset B200_OZ_SPLIT2
timeout 300 python profiles/host_profile.py 100 1024 > $T/r02o_host_profile.txt 2> $T/r02o_host_profile.err; head -3 $T/r02o_host_profile.txt; tail -c 300 $T/r02o_host_profile.err
timeout 900 python bench.py --steps 1 --warmup 3 --no-cpu > $T/r02o_bench.json 2> $T/r02o_bench.err; tail -c 300 $T/r02o_bench.err
python -c "
import json; d=json.load(open('$T/r02o_bench.json')); print(d['value'], d['e2e']['value'], d['kernel_family_ms_per_sweep'], d['roofline']['achieved'], d['parity'])"
