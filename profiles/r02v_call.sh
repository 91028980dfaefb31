#!/bin/bash
# round 2, call V: q = 4096 fix of the reorder kernel, chi scan on one GPU
T=gpurun_out
mkdir -p $T
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "svd" > $T/r02v_tests.log 2>&1; tail -n 3 $T/r02v_tests.log
timeout 1200 python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --no-blocksparse --scan on > $T/r02v_bench_scan.json 2> $T/r02v_bench_scan.err; tail -c 300 $T/r02v_bench_scan.err
python -c "
import json; d=json.load(open('$T/r02v_bench_scan.json')); print(d['value'], json.dumps(d['chi_scan'])[:1800])"
