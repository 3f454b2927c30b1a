#!/bin/bash
# round 2, call l (1 GPU): banded host-image step on hardware, default bench line
O=gpurun_out/r2l; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -6 > $O/tests.log
( time python bench.py ) > $O/bench_default.json 2> $O/bench_default.err
python bench.py --workload A --steps 50 --no-extras --no-cpu-baseline > $O/bench_A.json 2> $O/bench_A.err
ls -la $O
