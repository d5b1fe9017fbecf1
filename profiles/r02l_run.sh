#!/bin/bash
# round 2, call l: full GPU suite on the current code, compute-sanitizer memcheck of the DEFAULT path (third-generation
# kernels, single partition and a two-partition group), the degree-imbalanced rough structure timed at ~100k atoms
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 500 python -m pytest tests -m gpu -x -q > $O/r02l_pytest.txt 2>&1
tail -4 $O/r02l_pytest.txt
timeout 200 compute-sanitizer --tool memcheck --launch-timeout 0 python tests/_run_case.py /tmp/san.npz 2 > $O/r02l_memcheck_default.txt 2>&1
echo "memcheck rc=$?" >> $O/r02l_memcheck_default.txt
tail -3 $O/r02l_memcheck_default.txt
cat > /tmp/grp.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
from distmlip_b200.implementations.matgl import CHGNet_Dist, Potential_Dist
from distmlip_b200.structures import si_diamond
from tests._util import make_model
dm = CHGNet_Dist.from_existing(make_model()); dm.enable_distributed_mode([0, 0])
E, F, S, _ = Potential_Dist(model=dm)(si_diamond(4, nz=8, seed=3))
print("group E", E.item(), "Fmax", F.abs().max().item())
PY
timeout 300 compute-sanitizer --tool memcheck --launch-timeout 0 python /tmp/grp.py > $O/r02l_memcheck_group.txt 2>&1
echo "memcheck rc=$?" >> $O/r02l_memcheck_group.txt
tail -4 $O/r02l_memcheck_group.txt
timeout 200 python bench.py --rough-atoms 100000 --steps 10 --warmup 3 --no-cpu-baseline > $O/r02l_bench_rough100k.json 2> $O/r02l_bench_rough100k.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02l_bench_rough100k.json'))
print('rough', d['config']['atoms'], 'atoms', d['config']['edges_per_gpu'], 'edges: ms/step', round(d['ms_per_step'],2), 'atoms/s', round(d['value']), 'gather_ms', round(d['roofline']['kernel_ms'],3), d['parity']['net_force_max'])
PY
