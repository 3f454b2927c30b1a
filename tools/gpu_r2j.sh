#!/bin/bash
# round 2, call j (1 GPU): deposit with four register slots in flight + push with tensor segment copies (shipped build)
O=gpurun_out/r2j; mkdir -p $O
B="python bench.py --no-extras --no-e2e --no-cpu-baseline"
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -4 > $O/parity.log
$B > $O/D.json 2> $O/D.err
$B --workload A --steps 50 > $O/A.json 2> $O/A.err
$B --workload cyc > $O/cyc.json 2> $O/cyc.err
ls -la $O
