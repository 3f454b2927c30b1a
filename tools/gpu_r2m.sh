#!/bin/bash
# round 2, call m (8 GPUs): final multi-rank evidence of HEAD: 4- and 8-rank parity, scaling lines at 8 GPUs
O=gpurun_out/r2m; mkdir -p $O
git_head=$(cat .git_head 2>/dev/null)
echo "HEAD $git_head" > $O/head.txt
timeout 700 python -m pytest tests/test_gpu_multi.py -m gpu -q -x -k "eight_ranks or four_ranks" 2>&1 | tail -15 > $O/multi_4_8.log
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
( time $T --nproc-per-node 8 --master-port 29701 bench.py --gpus 8 ) > $O/bench8_default.json 2> $O/bench8_default.err
$T --nproc-per-node 8 --master-port 29702 bench.py --gpus 8 --workload A --steps 30 --warmup 5 --no-extras --no-e2e > $O/bench8_A.json 2> $O/bench8_A.err
ls -la $O
