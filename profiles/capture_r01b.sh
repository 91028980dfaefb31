#!/bin/bash
# Round-1 follow-up capture for the split Jacobi kernels (gram / eig / apply): shorter chain so that the run under
# ncu stays within a few minutes; the kernels and the per-bond shapes are those of the L=100 benchmark.
set -x
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 30000 --csv \
    --log-file gpurun_out/r01b_launches.csv python bench.py --L 40 --steps 1 --warmup 1 --no-e2e --no-cpu \
    > gpurun_out/r01b_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:jacobi_ -s 90 -c 6 \
    -o gpurun_out/r01b_jacobi python tests/dev_kernel_probe.py svd > gpurun_out/r01b_jacobi.log 2>&1
