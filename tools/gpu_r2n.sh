#!/bin/bash
# round 2, call n (2 GPUs): final 2-rank evidence of HEAD: parity over peer memory and over NCCL, default line
O=gpurun_out/r2n; mkdir -p $O
cat .git_head > $O/head.txt
timeout 900 python -m pytest tests/test_gpu_multi.py "tests/test_gpu_zz_robustness.py::test_two_ranks_streamed_initialisation" -m gpu -q -x 2>&1 | tail -15 > $O/multi2.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701"
( time $T bench.py --gpus 2 ) > $O/bench2_default.json 2> $O/bench2_default.err
ls -la $O
