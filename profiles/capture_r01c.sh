#!/bin/bash
# Round-1 capture (c): launch list + ncu --set full of the kernels of ONE centre-bond update at chi=1024
# (profiles/bond_probe.py); 64x64 GEMM tiles, split Jacobi kernels.  ~2-3 minutes on the box.
set -x
T=gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $T/r01c_launches.csv python profiles/bond_probe.py --bonds 2 > $T/r01c_launches_probe.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:grouped_gemm -c 4 \
    -o $T/r01c_gemm -f python profiles/bond_probe.py --bonds 1 > $T/r01c_gemm.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:jacobi_ -s 60 -c 6 \
    -o $T/r01c_jacobi -f python profiles/bond_probe.py --bonds 1 > $T/r01c_jacobi.log 2>&1
tail -1 $T/r01c_launches_probe.log $T/r01c_gemm.log $T/r01c_jacobi.log
