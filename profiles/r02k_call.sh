#!/bin/bash
# round 2, call K: per-phase timing of a bond update (identity shortcut on / off), ncu captures of the current kernels on
# one centre-bond update of the benchmark (launch list + --set full of the first launches of every kernel)
T=gpurun_out
mkdir -p $T
timeout 300 python profiles/bond_phases.py 30 1024 > $T/r02k_bond_phases.jsonl 2> $T/r02k_bond_phases.err; cat $T/r02k_bond_phases.jsonl; tail -c 300 $T/r02k_bond_phases.err
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $T/r02k_launches.csv python profiles/bond_probe.py --bonds 1 > $T/r02k_launches_probe.log 2>&1; tail -n 1 $T/r02k_launches_probe.log
for pat in 'oz_gemm' 'oz_split|oz_rowmax|mid_contract2' 'jacobi_gram|jacobi_eig|jacobi_apply' 'copy_blocks|take_blocks|lanczos_update|dot_partial|axpy_kernel|scal' 'grouped_gemm|thin_'; do
  tag=$(echo $pat | tr -c 'a-z0-9_' '_' | cut -c1-20)
  timeout 200 ncu --set full --clock-control none --profile-from-start off -k regex:"$pat" -c 6 \
      -o $T/r02k_$tag -f python profiles/bond_probe.py --bonds 1 > $T/r02k_ncu_$tag.log 2>&1; tail -n 1 $T/r02k_ncu_$tag.log
done
ls -la $T | grep r02k
