mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r1c.txt 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_r1c.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r1c.txt 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_r1c.txt
timeout 200 python bench.py > gpurun_out/bench_r1c.json 2> gpurun_out/bench_r1c.err; echo "bench rc=$?"; cut -c1-330 gpurun_out/bench_r1c.json; grep -o '"clocks": {[^}]*}' gpurun_out/bench_r1c.json
timeout 240 ncu --set full --clock-control none --import-source on -k regex:'k_atomconv_(fwd|bwd)_tc|k_line_(fwd|bwd)_tc' --launch-skip 3 --launch-count 10 -f -o gpurun_out/prof_r1c python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_r1c.log 2>&1; echo "ncu rc=$?"; ls -la gpurun_out/prof_r1c.ncu-rep
