#!/bin/bash
# round 2, call e (8 GPUs): 4- and 8-rank parity over peer memory, scaling lines at 8 and 4 GPUs
O=gpurun_out/r2e; mkdir -p $O
nvidia-smi -L > $O/gpus.txt
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -x -k "eight_ranks or (four_ranks and far)" 2>&1 | tail -15 > $O/multi_4_8.log
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
( time $T --nproc-per-node 8 --master-port 29701 bench.py --gpus 8 ) > $O/bench8_default.json 2> $O/bench8_default.err
$T --nproc-per-node 8 --master-port 29702 bench.py --gpus 8 --workload A --steps 30 --warmup 5 --no-extras --no-e2e > $O/bench8_A.json 2> $O/bench8_A.err
( time $T --nproc-per-node 4 --master-port 29703 bench.py --gpus 4 --no-e2e ) > $O/bench4_default.json 2> $O/bench4_default.err
ls -la $O
