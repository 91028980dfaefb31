#!/bin/bash
# round 2, call W: parity suites after the correction of the completion rule; ncu evidence of the final kernels on one
# centre-bond update of the benchmark at chi = 1024
T=gpurun_out
mkdir -p $T
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_large_parity.py tests/test_tebd.py tests/test_qr_truncation.py tests/test_dropin_engine.py tests/test_single_site.py -m gpu -q > $T/r02w_tests.log 2>&1; tail -n 3 $T/r02w_tests.log
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $T/r02w_launches.csv python profiles/bond_probe.py --bonds 1 > $T/r02w_launches_probe.log 2>&1; tail -n 1 $T/r02w_launches_probe.log
for pat in 'oz_gemm' 'oz_split' 'jacobi_' ; do
  tag=$(echo $pat | tr -c 'a-z0-9_' '_' | cut -c1-20)
  timeout 200 ncu --set full --clock-control none --profile-from-start off -k regex:"$pat" -c 6 \
      -o $T/r02w_$tag -f python profiles/bond_probe.py --bonds 1 > $T/r02w_ncu_$tag.log 2>&1; tail -n 1 $T/r02w_ncu_$tag.log
done
ls -la $T | grep r02w
