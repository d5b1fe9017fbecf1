#!/bin/bash
# round 2, call e: third-generation backward (k_atomconv_bwd_v3) first run + forward with L1 prefetch of C rows
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
B2M_TEST_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gpu_experimental.py -m gpu -q -k "ATOMCONV" > $O/r02e_experimental.txt 2>&1
tail -5 $O/r02e_experimental.txt
for gen in 4 3; do
B2M_ATOMCONV=$gen timeout 300 python bench.py --cells 23 --steps 10 --warmup 3 --no-cpu-baseline > $O/r02e_bench_97k_gen$gen.json 2> $O/r02e_bench_97k_gen$gen.err
python - <<PY
import json
d=json.load(open('gpurun_out/r02e_bench_97k_gen$gen.json'))
print('gen$gen ms/step', d['ms_per_step'], d['phase_ms'], 'gather_ms', d['roofline']['kernel_ms'], d['parity'])
PY
done
B2M_ATOMCONV=3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_atomconv_bwd_v3 -s 1 -c 1 \
  -o $O/r02e_bwd_v3 python bench.py --cells 23 --steps 1 --warmup 0 --no-cpu-baseline > $O/r02e_ncu.log 2>&1
