#!/bin/bash
# round 2, call ae: final-code validation (CHGNet + TensorNet): GPU suite, smoke(), default bench line, TensorNet line + launch list
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 240 python -c "import torch; print(torch.cuda.get_device_name(0))"   # image page-in happens here, not inside a test timeout
timeout 300 python -m pytest tests -m gpu -x -q > $O/r02ae_pytest.txt 2>&1
tail -3 $O/r02ae_pytest.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02ae_smoke.txt 2>&1
tail -3 $O/r02ae_smoke.txt
timeout 200 python bench.py --steps 20 --warmup 5 > $O/r02ae_bench_1M_n1.json 2> $O/r02ae_bench_1M_n1.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02ae_bench_1M_n1.json') if l.startswith('{')][-1])
print('N=1 1M: ms/step', round(d['ms_per_step'],2), 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'frac', round(d['roofline']['frac'],4), 'gather', round(d['roofline']['kernel_ms'],3), d['clocks'])
PY
timeout 120 python bench.py --model tensornet --cells 23 --steps 10 --warmup 3 > $O/r02ae_bench_tn_97k.json 2> $O/r02ae_bench_tn_97k.err
cut -c1-330 $O/r02ae_bench_tn_97k.json
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r02ae_launches_tn_97k.csv \
  python bench.py --model tensornet --cells 23 --steps 1 --warmup 0 > $O/r02ae_ncu.log 2>&1
