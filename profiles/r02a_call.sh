#!/bin/bash
# round 2, call A: round-1 leftovers (opt-in kernels A/B) + configs[2]/[3] first end-to-end runs
bash profiles/round2_first_call.sh
T=gpurun_out
timeout 420 python profiles/blocksparse_dmrg_probe.py xxz --L 100 --chi 1024 --ramp 6 --timed 2 > $T/r02a_xxz.json 2> $T/r02a_xxz.err; tail -c 600 $T/r02a_xxz.err; cat $T/r02a_xxz.json
timeout 420 python profiles/blocksparse_dmrg_probe.py hubbard --L 32 --chi 1024 --ramp 6 --timed 2 > $T/r02a_hubbard.json 2> $T/r02a_hubbard.err; tail -c 600 $T/r02a_hubbard.err; cat $T/r02a_hubbard.json
