#!/bin/bash
# round 2, call B: tcgen05 int8 bring-up probe + where-does-the-time-go for configs[2] (XXZ)
T=gpurun_out
mkdir -p $T
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $T/r02b_smi.txt
timeout 120 ./profiles/tc_i8_probe all > $T/r02b_tc_probe.log 2>&1; cat $T/r02b_tc_probe.log
timeout 300 python profiles/blocksparse_dmrg_probe.py xxz --L 100 --chi 1024 --ramp 6 --timed 2 > $T/r02b_xxz.json 2> $T/r02b_xxz.err; tail -c 300 $T/r02b_xxz.err; python -c "
import json; d=json.loads(open('$T/r02b_xxz.json').read().strip().splitlines()[-1]); print(json.dumps({k: d[k] for k in ('timed_sweep_s','family_ms_last_sweep','detail')}, indent=1))"
