#!/bin/bash
# round 2, call m: tile-interleaved angle features + next-tile index / block prefetch in the line-graph kernels
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_group.py -m gpu -x -q > $O/r02m_pytest.txt 2>&1
tail -4 $O/r02m_pytest.txt
timeout 100 python bench.py --cells 23 --steps 10 --warmup 3 --no-cpu-baseline > $O/r02m_bench_97k.json 2> $O/r02m_bench_97k.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02m_bench_97k.json'))
print('ms/step', round(d['ms_per_step'],3), d['phase_ms'], 'E/atom', d['parity']['energy_per_atom'])
PY
timeout 150 ncu --metrics gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:k_line_ -s 4 -c 10 --csv --log-file $O/r02m_line_times.csv python bench.py --cells 23 --steps 1 --warmup 0 --no-cpu-baseline > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02m_line_times.csv')) if len(r)>5]
h=rows[0]; k=h.index('Kernel Name'); m=h.index('Metric Name'); v=h.index('Metric Value')
for r in rows[1:]:
    print(r[k][:28], r[m][:40], r[v])
PY
