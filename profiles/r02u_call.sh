#!/bin/bash
# round 2, call U (final): the whole GPU suite, the contract's bench lines (both arms), smoke(), ncu evidence of the current
# kernels (launch list of one centre-bond update + --set full of the first launches of the dominant kernels)
T=gpurun_out
mkdir -p $T
timeout 1500 python -m pytest tests -m gpu -q > $T/r02u_tests.log 2>&1; tail -n 4 $T/r02u_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $T/r02u_smoke.log 2>&1; tail -n 2 $T/r02u_smoke.log
timeout 900 python bench.py --steps 1 --warmup 3 > $T/r02u_bench.json 2> $T/r02u_bench.err; tail -c 300 $T/r02u_bench.err
python -c "
import json; d=json.load(open('$T/r02u_bench.json')); print(d['value'], d['e2e']['value'], d['kernel_family_ms_per_sweep'], d['roofline']['achieved'], d['roofline']['frac'], d['parity'], d['cpu_baseline'], d['reference_default_lanczos'], d['reference_driver'].get('value'), d['clocks'])"
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 --ref-budget-s 90 > $T/r02u_bench_reference.json 2> $T/r02u_bench_reference.err; tail -c 200 $T/r02u_bench_reference.err; head -c 1200 $T/r02u_bench_reference.json
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $T/r02u_launches.csv python profiles/bond_probe.py --bonds 1 > $T/r02u_launches_probe.log 2>&1; tail -n 1 $T/r02u_launches_probe.log
for pat in 'oz_gemm' 'oz_split' 'jacobi_' ; do
  tag=$(echo $pat | tr -c 'a-z0-9_' '_' | cut -c1-20)
  timeout 200 ncu --set full --clock-control none --profile-from-start off -k regex:"$pat" -c 6 \
      -o $T/r02u_$tag -f python profiles/bond_probe.py --bonds 1 > $T/r02u_ncu_$tag.log 2>&1; tail -n 1 $T/r02u_ncu_$tag.log
done
timeout 300 python profiles/timeline_gaps.py 40 1024 4 > $T/r02u_timeline.json 2> $T/r02u_timeline.err
ls -la $T | grep r02u
