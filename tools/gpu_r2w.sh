#!/bin/bash
# round 2, call w (1 GPU): size-independent properties at the default workload's full size
O=gpurun_out/r2w; mkdir -p $O
( time timeout 200 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -x -k "config_D" ) > $O/fullsize_D.log 2>&1
tail -8 $O/fullsize_D.log
