#!/bin/bash
# round 2, call j (2 GPUs): NCCL multi-rank parity on final code (forward halo exchange on the second stream), single-process
# group on two distinct devices (peer-memory halo stores over NVLink), bench on the metric's 1M-atom cell at N=1 and N=2
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > $O/r02j_gpus.txt
timeout 200 python -m pytest tests/test_gpu_multirank.py tests/test_gpu_group.py tests/test_gpu_ase.py -m gpu -q > $O/r02j_multi.txt 2>&1
tail -6 $O/r02j_multi.txt
timeout 240 python bench.py > $O/r02j_bench_1M_n1.json 2> $O/r02j_bench_1M_n1.err
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > $O/r02j_bench_1M_n2.json 2> $O/r02j_bench_1M_n2.err
python - <<'PY'
import json
for n in (1,2):
    try:
        d=json.loads([l for l in open(f'gpurun_out/r02j_bench_1M_n{n}.json') if l.startswith('{')][-1])
        print(n, 'ms/step', round(d['ms_per_step'],2), 'value', round(d['value']), 'e2e', round(d['e2e']['value']), d['phase_ms'], 'frac', round(d['roofline']['frac'],4))
        print('   parity', d['parity'])
    except Exception as ex:
        print(n, 'FAILED', ex)
PY
tail -3 $O/r02j_bench_1M_n2.err
