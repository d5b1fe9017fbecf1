#!/bin/bash
# round 2, call q (2 GPUs): does the NCCL p2p channel count explain the 7 ms of exchange time at N=2 (r02j: 107.8 ms vs 100.4 ms
# for one half alone)?  bench N=2 on the 1M-atom cell with NCCL defaults and with NCCL_MIN_P2P_NCHANNELS=16 (new default)
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
for ch in 0 16; do
B2M_NCCL_P2P_CHANNELS=$ch timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2953$((ch%10)) bench.py --gpus 2 --no-cpu-baseline > $O/r02q_bench_1M_n2_ch$ch.json 2> $O/r02q_bench_1M_n2_ch$ch.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r02q_bench_1M_n2_ch$ch.json') if l.startswith('{')][-1])
print('channels $ch: N=2 ms/step', round(d['ms_per_step'],2), 'value', round(d['value']), 'e2e', round(d['e2e']['value']), d['phase_ms'], d['parity'].get('vs_single_partition'))
PY
done
