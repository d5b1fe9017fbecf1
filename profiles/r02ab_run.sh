#!/bin/bash
# round 2, call ab: TensorNet bench line at 97k atoms + launch list (kernel shares of a step)
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 300 python bench.py --model tensornet --cells 23 --steps 5 --warmup 3 > $O/r02ab_bench_tn_97k.json 2> $O/r02ab_bench_tn_97k.err
tail -c 1500 $O/r02ab_bench_tn_97k.json; tail -3 $O/r02ab_bench_tn_97k.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02ab_launches_tn_97k.csv \
  python bench.py --model tensornet --cells 23 --steps 1 --warmup 0 > $O/r02ab_ncu.log 2>&1
tail -2 $O/r02ab_ncu.log
