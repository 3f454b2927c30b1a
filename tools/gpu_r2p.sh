#!/bin/bash
# round 2, call p (1 GPU): compute-sanitizer over the parity suite (memcheck) and over a fused run with arrivals,
# leavers and far movers (racecheck: shared-memory hazards of the push's scratch and the deposit's turn taking)
O=gpurun_out/r2p; mkdir -p $O
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "first_10_steps or far_movers or hot_beam or deposit or banded or host_memory or device_init" > $O/memcheck.log 2>&1
echo "memcheck rc=$?" >> $O/memcheck.log
timeout 900 compute-sanitizer --tool racecheck --racecheck-report all --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "far_movers or hot_beam or deposit_parity or (first_10_steps and 2species)" > $O/racecheck.log 2>&1
echo "racecheck rc=$?" >> $O/racecheck.log
tail -5 $O/memcheck.log $O/racecheck.log
