#!/bin/bash
# round 2, call o (8 GPUs): the metric's 1M-atom cell at N=8 (strong scaling, NCCL transport) with the in-run parity block,
# 8-rank parity vs the oracle, and the single-process group over 8 devices (peer stores) at the same size
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 tests/run_multirank.py > $O/r02o_multirank8.txt 2>&1
grep -E "world|MULTIRANK" $O/r02o_multirank8.txt | tail -4
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 > $O/r02o_bench_1M_n8.json 2> $O/r02o_bench_1M_n8.err
python - <<'PY'
import json
try:
    d=json.loads([l for l in open('gpurun_out/r02o_bench_1M_n8.json') if l.startswith('{')][-1])
    print('N=8 ms/step', round(d['ms_per_step'],2), 'value', round(d['value']), 'e2e', round(d['e2e']['value']), d['phase_ms'])
    print('   parity', d['parity'])
except Exception as ex:
    print('N=8 FAILED', ex)
PY
tail -2 $O/r02o_bench_1M_n8.err
cat > /tmp/grp8.py <<'PY'
import sys, os, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from distmlip_b200.implementations.matgl import CHGNet_Dist, Potential_Dist
from distmlip_b200.structures import si_diamond
from distmlip_b200.random_init import RandomCHGNet
atoms = si_diamond(50)
dm = CHGNet_Dist.from_existing(RandomCHGNet(seed=0)); dm.enable_distributed_mode(list(range(8)))
pot = Potential_Dist(model=dm)
out = pot(atoms)
for _ in range(2): dm._engine.compute_resident(1)
ts=[]
for _ in range(5):
    e, ms = dm._engine.compute_resident(1); ts.append(ms)
t0=time.perf_counter()
for _ in range(3): out = pot(atoms)
e2e=(time.perf_counter()-t0)/3
F = out[1].numpy()
print("group8 1M atoms: device ms/step", np.mean(ts), "atoms/s", 1e6/np.mean(ts)*1e3, "e2e ms", e2e*1e3, "E/atom", out[0].item()/1e6, "netF", np.abs(F.sum(0)).max())
PY
timeout 200 python /tmp/grp8.py > $O/r02o_group8.txt 2>&1
tail -2 $O/r02o_group8.txt
