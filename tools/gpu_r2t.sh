#!/bin/bash
# round 2, call t (1 GPU): the tests added after the last full run
O=gpurun_out/r2t; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "host_order or banded or host_memory" 2>&1 | tail -5 > $O/tests.log
cat $O/tests.log
