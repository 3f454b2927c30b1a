#!/bin/bash
# round 2, call s (8 GPUs): 8-rank (and one 4-rank) parity of the final HEAD
O=gpurun_out/r2s; mkdir -p $O
cat .git_head > $O/head.txt
timeout 500 python -m pytest tests/test_gpu_multi.py -m gpu -q -x -k "eight_ranks or (four_ranks and far)" 2>&1 | tail -8 > $O/multi_4_8.log
cat $O/multi_4_8.log
