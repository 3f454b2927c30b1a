#!/bin/bash
# round 2, call u (1 GPU): the whole -m gpu suite of the final HEAD
O=gpurun_out/r2u; mkdir -p $O
cat .git_head > $O/head.txt
( time timeout 420 python -m pytest tests/ -x -q -m gpu ) > $O/gpu_suite.log 2>&1
tail -6 $O/gpu_suite.log
