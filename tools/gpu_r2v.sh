#!/bin/bash
# round 2, call v (1 GPU): the library rebuilt from a clean tree: smoke + a slice of the parity suite
O=gpurun_out/r2v; mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "first_10_steps_fused or deposit or banded" 2>&1 | tail -3 >> $O/smoke.log
cat $O/smoke.log
