#!/bin/bash
# round 2, call ac: TensorNet edge MLP on the tcgen05 row GEMM (+ activation reverse folded into the message reverse)
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 400 python -m pytest tests/test_gpu_tensornet.py -q -m gpu 2>&1 | tail -60 > $O/r02ac_tn_tests.txt
B2M_TN_FFMA=1 timeout 200 python -m pytest tests/test_gpu_tensornet.py -q -m gpu -k "stage_taps or larger" 2>&1 | tail -30 > $O/r02ac_tn_tests_ffma.txt
timeout 300 python bench.py --model tensornet --cells 23 --steps 5 --warmup 3 > $O/r02ac_bench_tn_97k.json 2> $O/r02ac_bench_tn_97k.err
timeout 400 python -m pytest tests/test_gpu_parity.py -q -m gpu -x 2>&1 | tail -5 > $O/r02ac_chgnet_parity.txt
tail -3 $O/r02ac_tn_tests.txt $O/r02ac_tn_tests_ffma.txt $O/r02ac_chgnet_parity.txt; cut -c1-400 $O/r02ac_bench_tn_97k.json; tail -3 $O/r02ac_bench_tn_97k.err
