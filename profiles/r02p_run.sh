#!/bin/bash
# round 2, call p: per-partition compute of the 8-way (and 2-way) split of the 1M-atom cell, alone on one GPU, halo
# exchanges skipped -> how much of the N=8 step (28.75 ms) is per-partition work and how much is exchange / waiting
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
{
B2M_DEBUG_NO_HALO=1 timeout 120 python tests/slab_timing.py 8 3
B2M_DEBUG_NO_HALO=1 timeout 120 python tests/slab_timing.py 8 0
B2M_DEBUG_NO_HALO=1 timeout 120 python tests/slab_timing.py 2 1
B2M_DEBUG_NO_HALO=1 timeout 120 python tests/slab_timing.py 4 1
} > $O/r02p_slab_timing.txt 2>&1
cat $O/r02p_slab_timing.txt
B2M_DEBUG_NO_HALO=1 timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 110 --csv --log-file $O/r02p_launches_slab8.csv python tests/slab_timing.py 8 3 > /dev/null 2>&1
