#!/bin/bash
# round 2, call r (2 GPUs): after the same-host check in the peer-memory set-up: peer memory still chosen, parity, A line
O=gpurun_out/r2r; mkdir -p $O
cat .git_head > $O/head.txt
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x -k "two_ranks" 2>&1 | tail -8 > $O/multi2.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701"
$T bench.py --gpus 2 --workload A --steps 30 --warmup 5 --no-extras --no-e2e > $O/bench2_A.json 2> $O/bench2_A.err
cat $O/multi2.log; python -c "
import json; d=json.load(open('$O/bench2_A.json')); print(d['ms_per_step'], d['roofline']['stage_ms_per_step'])"
