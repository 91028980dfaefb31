#!/bin/bash
# round 2, call N: 8 epilogue warps in oz_gemm_kernel; where the GPU waits for the host (bond_busy); parity suites;
# inner sweeps of the pivot solver on the block-sparse workloads
T=gpurun_out
mkdir -p $T
B200_OZ_DEBUG=1 timeout 120 python profiles/ozaki_one.py 7 2 > $T/r02n_oz_debug.log 2>&1; tail -2 $T/r02n_oz_debug.log
timeout 200 python profiles/ozaki_bench.py 1024 > $T/r02n_ozaki.jsonl 2> $T/r02n_ozaki.err
python - <<'PY'
import json
for line in open('gpurun_out/r02n_ozaki.jsonl'):
    d = json.loads(line)
    print(d['shape'], ' '.join('%s mm %.3f ms (%.0f Tops, %.1f TF) err %.1e' % (k, d[k]['mm_ms'], d[k]['int8_Tops'], d[k]['mm_fp64_equiv_tflops'], d[k]['max_abs_diff_vs_dmma_rel']) for k in ('s7', 's8', 's9')))
PY
timeout 900 python -m pytest tests/test_ozaki.py tests/test_large_parity.py tests/test_dropin_engine.py -m gpu -q > $T/r02n_tests.log 2>&1; tail -n 5 $T/r02n_tests.log
timeout 300 python profiles/bond_busy.py 100 1024 > $T/r02n_bond_busy.json 2> $T/r02n_bond_busy.err; tail -c 600 $T/r02n_bond_busy.err
for inner in 0 1; do
  timeout 600 python bench.py --workload xxz --steps 2 --warmup 1 --svd-inner-sweeps $inner > $T/r02n_xxz_in$inner.json 2> $T/r02n_xxz_in$inner.err
  python -c "
import json; d=json.load(open('$T/r02n_xxz_in$inner.json')); print('xxz inner', $inner, d['value'], d.get('kernel_family_ms_per_sweep'), d['result'])"
done
timeout 900 python bench.py --steps 1 --warmup 3 > $T/r02n_bench.json 2> $T/r02n_bench.err; tail -c 300 $T/r02n_bench.err
python -c "
import json; d=json.load(open('$T/r02n_bench.json')); print(d['value'], d['e2e'], d['kernel_family_ms_per_sweep'], d['roofline']['achieved'], d['roofline']['frac'], d.get('reference_driver'), d['reference_default_lanczos'])"
