#!/bin/bash
# round 2, call t: line-graph forward v3 (summed gathered rows staged during the tensor-core waits); also B2M_L2_PREFETCH=2 again
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
B2M_TEST_EXPERIMENTAL=1 timeout 200 python -m pytest tests/test_gpu_experimental.py -m gpu -q -k "LINEFWD" > $O/r02t_experimental.txt 2>&1
tail -3 $O/r02t_experimental.txt
timeout 250 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $O/r02t_pytest.txt 2>&1
tail -3 $O/r02t_pytest.txt
for v in "B2M_LINEFWD=3" "B2M_LINEFWD=1" "B2M_L2_PREFETCH=2"; do
env $v timeout 100 python bench.py --cells 23 --steps 10 --warmup 3 --no-cpu-baseline > $O/r02t_bench_97k_$v.json 2> /dev/null
python - <<PY
import json
d=json.load(open('gpurun_out/r02t_bench_97k_$v.json'))
print('$v ms/step', round(d['ms_per_step'],3), d['phase_ms'], 'E/atom', d['parity']['energy_per_atom'])
PY
done
timeout 150 ncu --metrics gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:k_line_fwd -s 0 -c 5 --csv --log-file $O/r02t_line_times.csv python bench.py --cells 23 --steps 1 --warmup 0 --no-cpu-baseline > /dev/null 2>&1
grep -E "k_line" $O/r02t_line_times.csv | awk -F'","' '{print $5, $(NF-2), $NF}' | cut -c1-120
