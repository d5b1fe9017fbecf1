#!/bin/bash
# round 2, call h: single-process multi-partition group ([0,0] / [0,0,0] on one GPU) vs the oracle; full GPU suite; smoke
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 240 python -m pytest tests/test_gpu_group.py -m gpu -x -q > $O/r02h_group.txt 2>&1
tail -15 $O/r02h_group.txt
timeout 400 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_group.py > $O/r02h_pytest.txt 2>&1
tail -4 $O/r02h_pytest.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02h_smoke.txt 2>&1
tail -3 $O/r02h_smoke.txt
