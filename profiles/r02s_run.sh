#!/bin/bash
# round 2, call s: small streaming kernels (bond node update fwd/bwd, angle init fwd/bwd) at 128 rows per block with 16-byte accesses
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_group.py -m gpu -x -q > $O/r02s_pytest.txt 2>&1
tail -3 $O/r02s_pytest.txt
timeout 100 python bench.py --cells 23 --steps 10 --warmup 3 --no-cpu-baseline > $O/r02s_bench_97k.json 2> /dev/null
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02s_bench_97k.json'))
print('ms/step', round(d['ms_per_step'],3), d['phase_ms'], 'E/atom', d['parity']['energy_per_atom'])
PY
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_bond_node|k_angle_init|k_edge_basis|k_edge_final|k_bond_init" -s 8 -c 16 --csv --log-file $O/r02s_small_times.csv python bench.py --cells 23 --steps 1 --warmup 0 --no-cpu-baseline > /dev/null 2>&1
grep -E "k_" $O/r02s_small_times.csv | awk -F'","' '{print $5, $NF}' | cut -c1-90
