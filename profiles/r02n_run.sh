#!/bin/bash
# round 2, call n: backward v3 reads the saved u|v once (both second-layer gradients from one pass)
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
B2M_TEST_EXPERIMENTAL=1 timeout 200 python -m pytest tests/test_gpu_experimental.py -m gpu -q -k "ATOMCONV" > $O/r02n_experimental.txt 2>&1
tail -3 $O/r02n_experimental.txt
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $O/r02n_pytest.txt 2>&1
tail -3 $O/r02n_pytest.txt
timeout 100 python bench.py --cells 23 --steps 10 --warmup 3 --no-cpu-baseline > $O/r02n_bench_97k.json 2> $O/r02n_bench_97k.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02n_bench_97k.json'))
print('ms/step', round(d['ms_per_step'],3), d['phase_ms'], 'E/atom', d['parity']['energy_per_atom'])
PY
timeout 150 ncu --metrics gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum --clock-control none -k regex:k_atomconv_bwd -s 1 -c 3 --csv --log-file $O/r02n_bwd_times.csv python bench.py --cells 23 --steps 1 --warmup 0 --no-cpu-baseline > /dev/null 2>&1
grep -E "k_atomconv" $O/r02n_bwd_times.csv | awk -F'","' '{print $5, $(NF-2), $NF}' | head -9
