#!/bin/bash
# round 2, call P: full GPU suite after the host-side changes (cached stream handle, grouped_gemm without a host sync, no
# completion of directions svd_min cuts), cross mode of the pivot solver, benchmark lines
T=gpurun_out
mkdir -p $T
timeout 1500 python -m pytest tests -m gpu -q -x > $T/r02p_tests.log 2>&1; tail -n 4 $T/r02p_tests.log
timeout 600 python profiles/svd_variants.py > $T/r02p_svd_variants.jsonl 2> $T/r02p_svd_variants.err; cat $T/r02p_svd_variants.jsonl | cut -c1-700; tail -c 300 $T/r02p_svd_variants.err
timeout 900 python bench.py --steps 1 --warmup 3 --no-cpu > $T/r02p_bench.json 2> $T/r02p_bench.err; tail -c 300 $T/r02p_bench.err
python -c "
import json; d=json.load(open('$T/r02p_bench.json')); print(d['value'], d['e2e']['value'], d['kernel_family_ms_per_sweep'], d['roofline']['achieved'], d['parity'], d['reference_default_lanczos'], d['reference_driver'].get('value'))"
timeout 600 python bench.py --workload xxz --steps 2 --warmup 1 > $T/r02p_xxz.json 2> $T/r02p_xxz.err
python -c "
import json; d=json.load(open('$T/r02p_xxz.json')); print('xxz', d['value'], d.get('kernel_family_ms_per_sweep'), d['result'], d['gemm_by_flops'])"
