#!/bin/bash
# round 2, call k: launch list of one step (97k atoms and 1M atoms) + ncu --set full of the default edge-gather kernel
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 130 -c 130 --csv --log-file $O/r02k_launches_97k.csv \
  python bench.py --cells 23 --steps 1 --warmup 1 --no-cpu-baseline > $O/r02k_l97.log 2>&1
timeout 250 ncu --set full --clock-control none --import-source on -k regex:k_atomconv_fwd_v3 -s 1 -c 1 \
  -o $O/r02k_fwd_v3_1M python bench.py --steps 1 --warmup 0 --no-cpu-baseline > $O/r02k_ncu1M.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_line_ -s 6 -c 4 \
  -o $O/r02k_line python bench.py --cells 23 --steps 1 --warmup 0 --no-cpu-baseline > $O/r02k_ncu_line.log 2>&1
ls -la $O/r02k_*
