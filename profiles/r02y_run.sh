#!/bin/bash
# round 2, call y: streaming loads for the per-launch streams (radial basis, indices) of the atom-conv kernels
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 250 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $O/r02y_pytest.txt 2>&1
tail -2 $O/r02y_pytest.txt
for i in 1 2; do
timeout 100 python bench.py --cells 23 --steps 20 --warmup 5 --no-cpu-baseline > $O/r02y_bench_97k_$i.json 2> /dev/null
python - <<PY
import json
d=json.load(open('gpurun_out/r02y_bench_97k_$i.json'))
print('ms/step', round(d['ms_per_step'],3), d['phase_ms'], 'gather', round(d['roofline']['kernel_ms'],4))
PY
done
