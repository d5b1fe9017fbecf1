#!/bin/bash
# round 2, call f: backward v3 with the tensor-core M.dbe product; forward v3 (scalar math) with / without L1 prefetch of C rows
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
B2M_TEST_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gpu_experimental.py -m gpu -q -k "ATOMCONV" > $O/r02f_experimental.txt 2>&1
tail -5 $O/r02f_experimental.txt
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --cells 23 --steps 10 --warmup 3 --no-cpu-baseline > $O/r02f_bench_$name.json 2> $O/r02f_bench_$name.err
  python - <<PY
import json
d=json.load(open('gpurun_out/r02f_bench_$name.json'))
print('$name ms/step', round(d['ms_per_step'],3), d['phase_ms'], 'gather_ms', round(d['roofline']['kernel_ms'],4), 'E/atom', d['parity']['energy_per_atom'])
PY
}
run gen3 B2M_ATOMCONV=3
run gen3_l1pf B2M_ATOMCONV=3 B2M_AC3_L1PF=1
run gen4 B2M_ATOMCONV=4
B2M_ATOMCONV=3 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_atomconv -c 16 --csv --log-file $O/r02f_kernel_times.csv python bench.py --cells 23 --steps 1 --warmup 0 --no-cpu-baseline > /dev/null 2>&1
grep -E "k_atomconv" $O/r02f_kernel_times.csv | awk -F, '{print $5, $NF}' | head -12
