#!/bin/bash
# round 2, call L: deferred identity check + recipe from cache (host time per bond), configs[3] parity with per-bond spectra
T=gpurun_out
mkdir -p $T
timeout 600 python -m pytest tests/test_large_parity.py tests/test_gpu_parity.py tests/test_interop_reference.py -m gpu -x -q > $T/r02l_tests.log 2>&1; tail -n 6 $T/r02l_tests.log
timeout 300 python profiles/bond_phases.py 30 1024 > $T/r02l_bond_phases.jsonl 2> $T/r02l_bond_phases.err; cat $T/r02l_bond_phases.jsonl; tail -c 300 $T/r02l_bond_phases.err
timeout 500 python bench.py --steps 1 --warmup 3 > $T/r02l_bench.json 2> $T/r02l_bench.err; tail -c 300 $T/r02l_bench.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02l_bench.json').read().strip().splitlines()[-1])
print('bench sweep_s', d['value'], 'e2e', d['e2e']['value'], 'E', d['result']['E'], d['kernel_family_ms_per_sweep'], 'launches', d['gpu_launches'])
print('ab', d['ab'], 'default lanczos', d['reference_default_lanczos']); print('reference_driver', d.get('reference_driver', {}).get('value'), d['identity_env_stats'])
PY
