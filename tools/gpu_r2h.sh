#!/bin/bash
# round 2, call h (1 GPU): new tests on hardware (host-image step, async output), default bench line
O=gpurun_out/r2h; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dropin.py tests/test_gpu_reference.py -x -q -m gpu 2>&1 | tail -6 > $O/tests.log
( time python bench.py ) > $O/bench_default.json 2> $O/bench_default.err
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
ls -la $O
