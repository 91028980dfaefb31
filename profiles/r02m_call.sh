#!/bin/bash
# round 2, call M: two accumulator sets (epilogue overlapping the next pass) vs one set of four; parity suites
T=gpurun_out
mkdir -p $T
for acc in 4 2; do
  B200_OZ_ACC=$acc B200_OZ_DEBUG=1 timeout 120 python profiles/ozaki_one.py 7 2 > $T/r02m_oz_debug_acc$acc.log 2>&1; tail -2 $T/r02m_oz_debug_acc$acc.log
  B200_OZ_ACC=$acc timeout 200 python profiles/ozaki_bench.py 1024 > $T/r02m_ozaki_acc$acc.jsonl 2> $T/r02m_ozaki_acc$acc.err
done
python - <<'PY'
import json
for acc in (4, 2):
    for line in open('gpurun_out/r02m_ozaki_acc%d.jsonl' % acc):
        d = json.loads(line)
        print('acc', acc, d['shape'], ' '.join('%s mm %.3f ms (%.0f Tops, %.1f TF) err %.1e' % (k, d[k]['mm_ms'], d[k]['int8_Tops'], d[k]['mm_fp64_equiv_tflops'], d[k]['max_abs_diff_vs_dmma_rel']) for k in ('s7', 's8', 's9')))
PY
timeout 900 python -m pytest tests/test_large_parity.py tests/test_dropin_engine.py tests/test_ozaki.py -m gpu -x -q > $T/r02m_tests.log 2>&1; tail -n 5 $T/r02m_tests.log
