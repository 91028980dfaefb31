#!/bin/bash
# round 2, call S: active-set bookkeeping of the block SVD on the device (vs host), Lanczos / bond update without host round
# trips before the SVD; parity suites; benchmark lines of the three workloads
T=gpurun_out
mkdir -p $T
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py tests/test_large_parity.py tests/test_single_site.py -m gpu -q > $T/r02s_tests.log 2>&1; tail -n 3 $T/r02s_tests.log
for host in 1 0; do
  if [ $host = 1 ]; then export B200_SVD_HOST_REORDER=1; else unset B200_SVD_HOST_REORDER; fi
  timeout 900 python bench.py --workload hubbard --steps 1 --warmup 1 > $T/r02s_hubbard_host$host.json 2> $T/r02s_hubbard_host$host.err
  python -c "
import json; d=json.load(open('$T/r02s_hubbard_host$host.json')); print('hubbard host_reorder', $host, d['value'], d.get('kernel_family_ms_per_sweep'), d['result'])"
  timeout 600 python bench.py --workload xxz --steps 2 --warmup 1 > $T/r02s_xxz_host$host.json 2> $T/r02s_xxz_host$host.err
  python -c "
import json; d=json.load(open('$T/r02s_xxz_host$host.json')); print('xxz host_reorder', $host, d['value'], d.get('kernel_family_ms_per_sweep'), d['result'])"
done
unset B200_SVD_HOST_REORDER
timeout 900 python bench.py --steps 1 --warmup 3 --no-cpu > $T/r02s_bench.json 2> $T/r02s_bench.err; tail -c 300 $T/r02s_bench.err
python -c "
import json; d=json.load(open('$T/r02s_bench.json')); print(d['value'], d['e2e']['value'], d['kernel_family_ms_per_sweep'], d['roofline']['achieved'], d['parity'], d['reference_default_lanczos'], d['reference_driver'].get('value'))"
timeout 300 python profiles/bond_busy.py 100 1024 > $T/r02s_bond_busy.json 2> $T/r02s_bond_busy.err
