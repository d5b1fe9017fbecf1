#!/bin/bash
# round 2, call z: ncu --set full summaries of the final default kernels (atom conv fwd/bwd v3, line fwd/bwd) at 97k atoms
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_atomconv_|k_line_" -s 4 -c 10 \
  -o $O/r02z_tile_kernels python bench.py --cells 23 --steps 1 --warmup 0 --no-cpu-baseline > $O/r02z_ncu.log 2>&1
ls -la $O/r02z_tile_kernels.ncu-rep
