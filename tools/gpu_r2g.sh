#!/bin/bash
# round 2, call g (1 GPU): push variants on workloads D and A
O=gpurun_out/r2g; mkdir -p $O
B="python bench.py --no-extras --no-e2e --no-cpu-baseline"
$B > $O/D_base.json 2> $O/D_base.err
$B --workload A --steps 50 > $O/A_base.json 2> $O/A_base.err
for v in C3 S S3 B0; do
  CPIC_B200_LIB=$PWD/cpic_b200/variants/libcpic_b200_$v.so $B > $O/D_$v.json 2> $O/D_$v.err
  CPIC_B200_LIB=$PWD/cpic_b200/variants/libcpic_b200_$v.so $B --workload A --steps 50 > $O/A_$v.json 2> $O/A_$v.err
done
ls -la $O
