#!/bin/bash
# round 2, call a: state of HEAD on the box (new large-size parity tests, 1M default bench), then the never-run
# round-1 experiment switches behind timeouts.
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $O/r02a_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > $O/r02a_pytest.txt 2>&1
timeout 300 python bench.py --cells 23 --steps 10 --warmup 3 --no-cpu-baseline > $O/r02a_bench_97k.json 2> $O/r02a_bench_97k.err
timeout 600 python bench.py > $O/r02a_bench_1M.json 2> $O/r02a_bench_1M.err
timeout 300 compute-sanitizer --tool memcheck --launch-timeout 0 python tests/_run_case.py /tmp/san.npz 2 > $O/r02a_sanitizer_memcheck_tc.txt 2>&1
echo "memcheck rc=$?" >> $O/r02a_sanitizer_memcheck_tc.txt
B2M_TEST_EXPERIMENTAL=1 timeout 900 python -m pytest tests/test_gpu_experimental.py -m gpu -q > $O/r02a_experimental.txt 2>&1
B2M_ATOMCONV_V2=1 timeout 300 python bench.py --cells 23 --steps 10 --warmup 3 --no-cpu-baseline > $O/r02a_bench_97k_v2.json 2> $O/r02a_bench_97k_v2.err
B2M_GEMM_PIPE=1 timeout 300 python bench.py --cells 23 --steps 10 --warmup 3 --no-cpu-baseline > $O/r02a_bench_97k_pipe.json 2> $O/r02a_bench_97k_pipe.err
tail -3 $O/r02a_pytest.txt; tail -5 $O/r02a_experimental.txt; cat $O/r02a_bench_97k.json | cut -c1-400
