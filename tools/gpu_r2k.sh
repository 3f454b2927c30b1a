#!/bin/bash
# round 2, call k (1 GPU): arrival pairs in the deposit; particle blocks of 16 x 16 cells against the default 8 x 8
O=gpurun_out/r2k; mkdir -p $O
B="python bench.py --no-extras --no-e2e --no-cpu-baseline"
$B > $O/D.json 2> $O/D.err
$B --workload A --steps 50 > $O/A.json 2> $O/A.err
CPIC_B200_BLOCK_CELLS=16 $B > $O/D_bc16.json 2> $O/D_bc16.err
CPIC_B200_BLOCK_CELLS=16 $B --workload A --steps 50 > $O/A_bc16.json 2> $O/A_bc16.err
CPIC_B200_BLOCK_CELLS=16 CPIC_B200_DEP_COLS=16 $B > $O/D_bc16c16.json 2> $O/D_bc16c16.err
CPIC_B200_BLOCK_CELLS=4 $B --workload A --steps 50 > $O/A_bc4.json 2> $O/A_bc4.err
ls -la $O
