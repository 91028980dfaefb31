#!/bin/bash
# First GPU call of round 2: the opt-in kernels written in round 1 after the GPU budget was spent (all host-checked, none
# GPU-run yet).  Each step has its own timeout; outputs under gpurun_out/r02a_*.
set -x
T=gpurun_out
mkdir -p $T
timeout 300 python -m pytest tests -m gpu -x -q > $T/r02a_gpu_tests.log 2>&1; tail -n 3 $T/r02a_gpu_tests.log
timeout 180 python tests/dev_eig_v2_gpu_check.py > $T/r02a_eig_v2.log 2>&1; tail -n 12 $T/r02a_eig_v2.log
timeout 180 python tests/dev_matvec_order_probe.py 1024 24 10 > $T/r02a_matvec_probe.log 2>&1; tail -n 5 $T/r02a_matvec_probe.log
timeout 300 python tests/dev_optins_gpu_check.py 1024 24 > $T/r02a_optins.log 2>&1; tail -n 6 $T/r02a_optins.log
# A/B of the whole benchmark step: default configuration against all opt-ins (tenpy_b200/optins.py)
timeout 200 python bench.py --steps 1 --warmup 3 --no-cpu > $T/r02a_bench_default.json 2> $T/r02a_bench_default.err; tail -c 400 $T/r02a_bench_default.err
B200_OPTINS=all timeout 200 python bench.py --steps 1 --warmup 3 --no-cpu > $T/r02a_bench_optins.json 2> $T/r02a_bench_optins.err; tail -c 400 $T/r02a_bench_optins.err
python - <<'PY'
import json
for tag in ('default', 'optins'):
    try:
        d = json.loads(open('gpurun_out/r02a_bench_%s.json' % tag).read().strip().splitlines()[-1])
        print(tag, 'sweep_s', d['value'], 'e2e', d['e2e']['value'], 'E', d['result']['E'], 'matvec', d['matvec_orders'])
    except Exception as e:
        print(tag, 'no result:', e)
PY
# configs[2] / [3] end to end (block-sparse DMRG at the target bond dimensions); separate call if the budget is tight
# timeout 600 python profiles/blocksparse_dmrg_probe.py xxz --L 100 --chi 1024 --ramp 6 --timed 2 > $T/r02a_xxz.json 2> $T/r02a_xxz.err
# timeout 900 python profiles/blocksparse_dmrg_probe.py hubbard --L 64 --chi 2048 --ramp 7 --timed 2 > $T/r02a_hubbard.json 2> $T/r02a_hubbard.err
