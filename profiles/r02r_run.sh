#!/bin/bash
# round 2, call r: library-owned pinned staging with threaded host copies -> e2e; full GPU suite; default bench (1M atoms, N=1);
# compute-sanitizer racecheck of the default path on the 64-atom case
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > $O/r02r_pytest.txt 2>&1
tail -3 $O/r02r_pytest.txt
timeout 300 python bench.py > $O/r02r_bench_1M_n1.json 2> $O/r02r_bench_1M_n1.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02r_bench_1M_n1.json') if l.startswith('{')][-1])
print('N=1 1M: ms/step', round(d['ms_per_step'],2), 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'e2e ms', round(d['e2e']['ms_per_step'],2), d['phase_ms'], 'frac', round(d['roofline']['frac'],4), d['clocks'])
print(d['cpu_baseline'])
PY
timeout 100 python bench.py --cells 23 --no-cpu-baseline > $O/r02r_bench_97k.json 2>/dev/null
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02r_bench_97k.json'))
print('97k: ms/step', round(d['ms_per_step'],2), 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'e2e ms', round(d['e2e']['ms_per_step'],2))
PY
timeout 240 compute-sanitizer --tool racecheck --launch-timeout 0 python tests/_run_case.py /tmp/san.npz 2 > $O/r02r_racecheck_default.txt 2>&1
echo "racecheck rc=$?" >> $O/r02r_racecheck_default.txt
tail -4 $O/r02r_racecheck_default.txt
