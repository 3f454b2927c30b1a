#!/bin/bash
# round 2, call o (1 GPU): final single-GPU evidence of HEAD: -m gpu suite, smoke, both bench arms, ncu captures
O=gpurun_out/r2o; mkdir -p $O
cat .git_head > $O/head.txt
( time timeout 1500 python -m pytest tests/ -x -q -m gpu --durations=5 ) > $O/gpu_suite.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
( time python bench.py --impl reference --steps 20 --warmup 5 ) > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err
( time python bench.py --steps 20 --warmup 5 ) > $O/bench_default.json 2> $O/bench_default.err
B="python bench.py --no-extras --no-e2e --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 80 --csv --log-file $O/launches_D.csv $B --steps 4 --warmup 3 > $O/under_ncu_D.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_gather_push|k_deposit|k_rho_assemble|k_phi_E|k_green" -s 12 -c 6 -o $O/prof_D $B --steps 4 --warmup 3 > $O/ncu_D.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_gather_push|k_deposit|k_rho_assemble|k_phi_E|k_green" -s 12 -c 6 -o $O/prof_A $B --workload A --steps 6 --warmup 3 > $O/ncu_A.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 100 --csv --log-file $O/launches_A.csv $B --workload A --steps 6 --warmup 3 > $O/under_ncu_A.log 2>&1
ls -la $O
