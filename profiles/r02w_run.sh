#!/bin/bash
# round 2, call w: streaming hints incl. the line backward's second u|v pass; bench at 97k and at the metric's 1M atoms
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 250 python -m pytest tests/test_gpu_parity.py tests/test_gpu_group.py -m gpu -x -q > $O/r02w_pytest.txt 2>&1
tail -2 $O/r02w_pytest.txt
timeout 100 python bench.py --cells 23 --steps 20 --warmup 5 --no-cpu-baseline > $O/r02w_bench_97k.json 2> /dev/null
timeout 300 python bench.py --steps 20 --warmup 5 > $O/r02w_bench_1M_n1.json 2> /dev/null
python - <<'PY'
import json
for f in ('r02w_bench_97k','r02w_bench_1M_n1'):
    d=json.loads([l for l in open(f'gpurun_out/{f}.json') if l.startswith('{')][-1])
    print(f, 'ms/step', round(d['ms_per_step'],3), 'value', round(d['value']), 'e2e', round(d['e2e']['value']), d['phase_ms'], 'gather', round(d['roofline']['kernel_ms'],4), 'frac', round(d['roofline']['frac'],4))
PY
