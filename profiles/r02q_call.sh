#!/bin/bash
# round 2, call Q (2 GPUs): the contract's N = 2 launch -- one process per GPU, NCCL broadcast / all-gather, chi scan (configs[4])
T=gpurun_out
mkdir -p $T
nvidia-smi --query-gpu=index,name --format=csv > $T/r02q_smi.txt
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus 2 --steps 1 --warmup 3 --no-cpu > $T/r02q_bench_n2.json 2> $T/r02q_bench_n2.err
tail -c 500 $T/r02q_bench_n2.err
python -c "
import json; d=json.loads(open('$T/r02q_bench_n2.json').read().strip().splitlines()[-1]); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e'], d['parity'], json.dumps(d['chi_scan'])[:1500])"
