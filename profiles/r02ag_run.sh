#!/bin/bash
# round 2, call ag: TensorNet tests on the final HEAD build
cd "$GRAFT_REPO_ROOT"
timeout 60 python -m pytest tests/test_gpu_tensornet.py -q -m gpu -x > gpurun_out/r02ag_pytest.txt 2>&1
tail -2 gpurun_out/r02ag_pytest.txt
