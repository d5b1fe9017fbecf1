#!/bin/bash
# round 2, call x: final-code validation (partition views, triclinic group test, streaming hints as committed)
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 450 python -m pytest tests -m gpu -x -q > $O/r02x_pytest.txt 2>&1
tail -3 $O/r02x_pytest.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02x_smoke.txt 2>&1
tail -2 $O/r02x_smoke.txt
timeout 300 python bench.py --steps 20 --warmup 5 > $O/r02x_bench_1M_n1.json 2> $O/r02x_bench_1M_n1.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02x_bench_1M_n1.json') if l.startswith('{')][-1])
print('N=1 1M: ms/step', round(d['ms_per_step'],2), 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'frac', round(d['roofline']['frac'],4), 'gather', round(d['roofline']['kernel_ms'],3), d['clocks'])
PY
