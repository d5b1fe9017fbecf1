mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r1b.txt 2>&1; echo "pytest default rc=$?"; tail -2 gpurun_out/pytest_r1b.txt
B2M_FWD_PREFETCH=1 B2M_BWD_PREFETCH=1 timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r1b_pf.txt 2>&1; echo "pytest pf rc=$?"; tail -2 gpurun_out/pytest_r1b_pf.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r1b.txt 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_r1b.txt
timeout 200 python bench.py > gpurun_out/bench_r1b_default.json 2> gpurun_out/bench_r1b_default.err; echo "bench rc=$?"; cat gpurun_out/bench_r1b_default.json | cut -c1-400
B2M_FWD_PREFETCH=1 B2M_BWD_PREFETCH=1 timeout 150 python bench.py --no-cpu-baseline > gpurun_out/bench_r1b_pf.json 2> gpurun_out/bench_r1b_pf.err; cut -c1-300 gpurun_out/bench_r1b_pf.json
B2M_L2_PREFETCH=0 timeout 150 python bench.py --no-cpu-baseline > gpurun_out/bench_r1b_nol2.json 2> /dev/null; cut -c1-300 gpurun_out/bench_r1b_nol2.json
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1b.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench_r1b.log 2>&1; echo "ncu rc=$?"
