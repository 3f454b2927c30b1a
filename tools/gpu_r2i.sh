#!/bin/bash
# round 2, call i (1 GPU): the push with one tensor copy per segment batch (variants/libdev.so) against the shipped build
O=gpurun_out/r2i; mkdir -p $O
B="python bench.py --no-extras --no-e2e --no-cpu-baseline"
CPIC_B200_LIB=$PWD/cpic_b200/variants/libdev.so timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -4 > $O/parity_dev.log
$B > $O/D_base.json 2> $O/D_base.err
CPIC_B200_LIB=$PWD/cpic_b200/variants/libdev.so $B > $O/D_dev.json 2> $O/D_dev.err
$B --workload A --steps 50 > $O/A_base.json 2> $O/A_base.err
CPIC_B200_LIB=$PWD/cpic_b200/variants/libdev.so $B --workload A --steps 50 > $O/A_dev.json 2> $O/A_dev.err
ls -la $O
