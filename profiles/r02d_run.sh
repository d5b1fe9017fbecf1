#!/bin/bash
# round 2, call d: v3 forward with packed FP32 pairs and hoisted C-row loads -- A/B parity, bench, ncu
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
B2M_TEST_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gpu_experimental.py -m gpu -q -k "ATOMCONV" > $O/r02d_experimental.txt 2>&1
tail -5 $O/r02d_experimental.txt
B2M_ATOMCONV=3 timeout 300 python bench.py --cells 23 --steps 10 --warmup 3 --no-cpu-baseline > $O/r02d_bench_97k_v3.json 2> $O/r02d_bench_97k_v3.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02d_bench_97k_v3.json'))
print('ms/step', d['ms_per_step'], d['phase_ms'], 'gather_ms', d['roofline']['kernel_ms'], d['parity'])
PY
B2M_ATOMCONV=3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_atomconv_fwd_v3 -s 1 -c 1 \
  -o $O/r02d_fwd_v3 python bench.py --cells 23 --steps 1 --warmup 0 --no-cpu-baseline > $O/r02d_ncu.log 2>&1
