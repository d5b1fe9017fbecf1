#!/bin/bash
# round 2, call u: final-code validation -- GPU suite, smoke, default bench line (1M atoms, N=1), 97k line, reference arm
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > $O/r02u_pytest.txt 2>&1
tail -3 $O/r02u_pytest.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02u_smoke.txt 2>&1
tail -2 $O/r02u_smoke.txt
timeout 300 python bench.py --steps 20 --warmup 5 > $O/r02u_bench_1M_n1.json 2> $O/r02u_bench_1M_n1.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02u_bench_1M_n1.json') if l.startswith('{')][-1])
print('N=1 1M: ms/step', round(d['ms_per_step'],2), 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'frac', round(d['roofline']['frac'],4), d['clocks'], d['cpu_baseline']['value'], d['cpu_baseline']['cores'])
PY
timeout 100 python bench.py --cells 23 --steps 20 --warmup 5 --no-cpu-baseline > $O/r02u_bench_97k.json 2>/dev/null
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02u_bench_97k.json'))
print('97k: ms/step', round(d['ms_per_step'],2), 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'frac', round(d['roofline']['frac'],4))
PY
timeout 200 python bench.py --impl reference --steps 5 --warmup 1 > $O/r02u_bench_reference.json 2>/dev/null
cut -c1-300 $O/r02u_bench_reference.json
