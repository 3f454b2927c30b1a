#!/bin/bash
# round 2, call d (2 GPUs): multi-rank parity over peer memory and over NCCL, 2-GPU bench lines
O=gpurun_out/r2d; mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py "tests/test_gpu_zz_robustness.py::test_two_ranks_streamed_initialisation" -m gpu -q -x 2>&1 | tail -15 > $O/multi_p2p.log
CPIC_B200_P2P=0 timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -x -k "two_ranks and (uniform or far)" 2>&1 | tail -15 > $O/multi_nccl.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701"
$T bench.py --gpus 2 --workload A --steps 30 --warmup 5 --no-extras --no-e2e > $O/bench2_A_p2p.json 2> $O/bench2_A_p2p.err
CPIC_B200_P2P=0 $T bench.py --gpus 2 --workload A --steps 30 --warmup 5 --no-extras --no-e2e > $O/bench2_A_nccl.json 2> $O/bench2_A_nccl.err
( time $T bench.py --gpus 2 ) > $O/bench2_default.json 2> $O/bench2_default.err
ls -la $O
