#!/bin/bash
# round 2, call i: group tests again (call h: two partition threads raced on the per-device smem-attribute guard; now locked)
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_group.py -m gpu -q > $O/r02i_group.txt 2>&1
tail -15 $O/r02i_group.txt
