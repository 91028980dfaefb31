#!/bin/bash
# round 2, call C: first run of the tcgen05 Ozaki GEMM (parity tests, then timing)
T=gpurun_out
mkdir -p $T
timeout 300 python -m pytest tests/test_ozaki.py -m gpu -x -q > $T/r02c_ozaki_tests.log 2>&1; tail -n 15 $T/r02c_ozaki_tests.log
timeout 200 python profiles/ozaki_bench.py 1024 > $T/r02c_ozaki_bench.jsonl 2> $T/r02c_ozaki_bench.err; tail -c 600 $T/r02c_ozaki_bench.err; cat $T/r02c_ozaki_bench.jsonl
