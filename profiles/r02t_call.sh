#!/bin/bash
# round 2, call T: kernel timeline of a converged sweep (where the GPU idles)
T=gpurun_out
mkdir -p $T
timeout 400 python profiles/timeline_gaps.py 40 1024 4 > $T/r02t_timeline.json 2> $T/r02t_timeline.err; tail -c 600 $T/r02t_timeline.err; head -c 1500 $T/r02t_timeline.json
