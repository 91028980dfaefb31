#!/bin/bash
# round 2, call J: parity suites on the GPU (configs[2]/[3] goldens, drop-in, interop), SVD warm-start A/B on the XXZ
# workload, ncu captures of the current kernels on one centre-bond update of the benchmark
T=gpurun_out
mkdir -p $T
timeout 900 python -m pytest tests/test_large_parity.py tests/test_interop_reference.py tests/test_dropin_engine.py -m gpu -x -q > $T/r02j_tests.log 2>&1; tail -n 6 $T/r02j_tests.log
for ws in off full default; do
  timeout 300 python bench.py --workload xxz --steps 1 --warmup 1 --svd-warm-start $ws > $T/r02j_xxz_$ws.json 2> $T/r02j_xxz_$ws.err; tail -c 200 $T/r02j_xxz_$ws.err
done
python - <<'PY'
import json
for ws in ('off', 'full', 'default'):
    try:
        d = json.loads(open('gpurun_out/r02j_xxz_%s.json' % ws).read().strip().splitlines()[-1])
        print('xxz warm start', ws, 'sweep_s', d['value'], d['kernel_family_ms_per_sweep'], d['result'])
    except Exception as e:
        print(ws, 'no result', e)
PY
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $T/r02j_launches.csv python profiles/bond_probe.py --bonds 1 > $T/r02j_launches_probe.log 2>&1; tail -n 1 $T/r02j_launches_probe.log
timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:'oz_gemm|oz_split|oz_rowmax|mid_contract2|jacobi_gram|jacobi_eig|jacobi_apply|copy_blocks|take_blocks|lanczos_update|grouped_gemm|dot_partial|axpy' -c 60 \
    -o $T/r02j_bond -f python profiles/bond_probe.py --bonds 1 > $T/r02j_bond.log 2>&1; tail -n 2 $T/r02j_bond.log
ls -la $T | grep r02j
