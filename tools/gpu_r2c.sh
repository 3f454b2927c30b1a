#!/bin/bash
# round 2, call c (1 GPU): deposit v3 (independent warps, node tiles) parity + timing, new bench.py end to end
O=gpurun_out/r2c; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -5 > $O/parity.log
python bench.py --workload A --steps 50 --warmup 5 --no-extras --no-e2e --no-cpu-baseline > $O/bench_A.json 2> $O/bench_A.err
( time python bench.py ) > $O/bench_default.json 2> $O/bench_default.err
ncu --set full --clock-control none --import-source on -k regex:"k_deposit|k_rho_assemble" -s 6 -c 2 -o $O/prof_dep \
    python bench.py --workload A --steps 6 --warmup 3 --no-extras --no-e2e --no-cpu-baseline > $O/ncu_dep.log 2>&1
ls -la $O
