#!/bin/bash
# round 2, call X (2 GPUs): the N = 2 launch of the contract with the chi scan (configs[4]) after the q = 4096 fix
T=gpurun_out
mkdir -p $T
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 \
    bench.py --gpus 2 --steps 1 --warmup 3 --no-cpu --no-e2e --no-blocksparse > $T/r02x_bench_n2.json 2> $T/r02x_bench_n2.err
tail -c 300 $T/r02x_bench_n2.err
python -c "
import json; d=json.loads(open('$T/r02x_bench_n2.json').read().strip().splitlines()[-1]); print(d['n_gpus'], d['value'], d['ms_per_step'], d['parity'], json.dumps(d['chi_scan'])[:1800])"
