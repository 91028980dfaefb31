#!/bin/bash
# Round-1 capture (d): ONE centre-bond update at chi=1024 (profiles/bond_probe.py, 3 warm-up sweeps as bench.py) with the
# 'split' matvec order and interned layouts: launch list, then ncu --set full of the first GEMM / Jacobi launches.
set -x
T=gpurun_out
mkdir -p $T
timeout 80 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $T/r01d_launches.csv python profiles/bond_probe.py --bonds 1 > $T/r01d_launches_probe.log 2>&1
timeout 100 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:'grouped_gemm|jacobi_gram|jacobi_eig|jacobi_apply' -c 44 \
    -o $T/r01d_bond -f python profiles/bond_probe.py --bonds 1 > $T/r01d_bond.log 2>&1
tail -n 1 $T/r01d_launches_probe.log
tail -n 1 $T/r01d_bond.log
ls -la $T
