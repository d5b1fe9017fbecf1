#!/bin/bash
# round 2, call aa: first GPU run of the TensorNet path (stage taps vs the mirror, E/F/stress vs the oracle)
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_tensornet.py -q -m gpu -k "stage_taps" 2>&1 | tail -150 > $O/r02aa_taps.txt
timeout 400 python -m pytest tests/test_gpu_tensornet.py -q -m gpu -k "not stage_taps" 2>&1 | tail -80 > $O/r02aa_rest.txt
tail -5 $O/r02aa_taps.txt; tail -5 $O/r02aa_rest.txt
