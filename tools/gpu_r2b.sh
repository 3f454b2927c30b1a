#!/bin/bash
# round 2, call b (1 GPU): parity of the new deposit on hardware, config A bench, variants, ncu
O=gpurun_out/r2b; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -5 > $O/parity.log
python bench.py --steps 50 --warmup 5 > $O/bench_A.json 2> $O/bench_A.err
B="python bench.py --steps 50 --warmup 5 --no-e2e --no-cpu-baseline"
for v in S P2 P8 D3; do
  CPIC_B200_LIB=$PWD/cpic_b200/variants/libcpic_b200_$v.so $B > $O/bench_A_$v.json 2> $O/bench_A_$v.err
done
CPIC_B200_DEP_COLS=8 $B > $O/bench_A_cols8.json 2> $O/bench_A_cols8.err
CPIC_B200_DEP_COLS=4 $B > $O/bench_A_cols4.json 2> $O/bench_A_cols4.err
$B --workload cyc --steps 20 > $O/bench_cyc.json 2> $O/bench_cyc.err
# launch list + full capture of the two hot kernels (never a bench value)
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 120 --csv --log-file $O/launches.csv \
    python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu-baseline > $O/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_gather_push|k_deposit" -s 9 -c 3 -o $O/prof \
    python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu-baseline > $O/ncu_full.log 2>&1
ls -la $O
