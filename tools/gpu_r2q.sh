#!/bin/bash
# round 2, call q (2 GPUs): the native driver on two ranks
O=gpurun_out/r2q; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -x -k "own_driver or over_nccl" 2>&1 | tail -15 > $O/driver2.log
cat $O/driver2.log
