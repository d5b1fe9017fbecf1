mkdir -p gpurun_out
timeout 100 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r1d.txt 2>&1; echo "pytest rc=$?"; tail -1 gpurun_out/pytest_r1d.txt
timeout 70 python bench.py > gpurun_out/bench_r1d.json 2> gpurun_out/bench_r1d.err; echo "bench rc=$?"; cut -c1-330 gpurun_out/bench_r1d.json
timeout 30 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r1d.txt 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_r1d.txt
