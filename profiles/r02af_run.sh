#!/bin/bash
# round 2, call af: TensorNet tests on HEAD (reverse edge MLP writing into the forward scratch buffers)
cd "$GRAFT_REPO_ROOT"
timeout 150 python -m pytest tests/test_gpu_tensornet.py -q -m gpu -x > gpurun_out/r02af_pytest.txt 2>&1
tail -3 gpurun_out/r02af_pytest.txt
