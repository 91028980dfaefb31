#!/bin/bash
# round 2, call F: coalesced int8 epilogue, fast rotation parameters, drop-in on the GPU, launch list of one block SVD
T=gpurun_out
mkdir -p $T
B200_OZ_DEBUG=1 timeout 120 python profiles/ozaki_one.py 7 2 > $T/r02f_oz_debug.log 2>&1; cat $T/r02f_oz_debug.log
timeout 200 python profiles/ozaki_bench.py 1024 > $T/r02f_ozaki.jsonl 2> $T/r02f_ozaki.err
python - <<'PY'
import json
for line in open('gpurun_out/r02f_ozaki.jsonl'):
    d = json.loads(line)
    print(d['shape'], 'dmma %.3f ms' % d['dmma_ms'], ' '.join('%s mm %.3f ms (%.0f Tops, %.1f TF) split %.3f+%.3f' % (k, d[k]['mm_ms'], d[k]['int8_Tops'], d[k]['mm_fp64_equiv_tflops'], d[k]['split_A_ms'], d[k]['split_B_ms']) for k in ('s7', 's8', 's9')))
PY
timeout 600 python -m pytest tests -m gpu -x -q > $T/r02f_tests.log 2>&1; tail -n 6 $T/r02f_tests.log
timeout 300 python profiles/svd_variants.py > $T/r02f_svd_variants.jsonl 2> $T/r02f_svd_variants.err; cat $T/r02f_svd_variants.jsonl; tail -c 300 $T/r02f_svd_variants.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:jacobi -s 300 -c 600 --csv --log-file $T/r02f_svd_launches.csv python profiles/svd_one.py 1024 3 > $T/r02f_svd_ncu.log 2>&1; tail -2 $T/r02f_svd_ncu.log
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/r02f_svd_launches.csv')) if len(r) > 10]
hdr = rows[0]; ik = hdr.index('Kernel Name'); iv = hdr.index('Metric Value')
agg = collections.defaultdict(list)
for r in rows[1:]:
    try: agg[r[ik].split('(')[0]].append(float(r[iv].replace(',', '')))
    except Exception: pass
for k, v in agg.items(): print(k, 'n=%d' % len(v), 'mean %.1f us' % (sum(v) / len(v) / 1e3), 'max %.1f' % (max(v) / 1e3))
PY
timeout 200 python bench.py --steps 1 --warmup 3 --no-cpu > $T/r02f_bench.json 2> $T/r02f_bench.err; tail -c 300 $T/r02f_bench.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02f_bench.json').read().strip().splitlines()[-1])
print('bench sweep_s', d['value'], 'e2e', d['e2e']['value'], 'E', d['result']['E'], d['kernel_family_ms_per_sweep'])
PY
