#!/bin/bash
# (probe binaries: python -m distmlip_b200.build --probes)
# round 2, call b: hardware facts behind the third-generation edge-gather kernels.
#  (1) tcgen05 TF32 with an MN-major B operand over the canonical K-major image of W (-> W^T without a second copy)
#  (2) gather staging throughput per SM: bulk copies vs per-lane LDG vs coalesced LDG+STS vs cp.async; MUFU rates
#  (3) ncu --set full of the (round-1, never profiled) bulk-copy forward k_atomconv_fwd_v2
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
B=distmlip_b200/csrc/build
{
  echo "== umma_probe: K-major baseline (TS, split)"; timeout 60 $B/umma_probe 1 0 1
  for cfg in "128 1024 128" "1024 128 128" "128 1024 1024" "1024 128 1024" "128 1024 256" "1024 128 2048"; do
    echo "== umma_probe MN-major B: lbo sbo kstep = $cfg"; timeout 60 $B/umma_probe 1 0 1 1 $cfg
  done
} > $O/r02b_umma_probe.txt 2>&1
{
  for m in 0 1 2 3 4 5 6; do timeout 60 $B/gather_probe $m 144; done
} > $O/r02b_gather_probe.txt 2>&1
B2M_ATOMCONV_V2=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_atomconv_fwd_v2 -s 1 -c 1 \
  -o $O/r02b_fwd_v2 python bench.py --cells 23 --steps 1 --warmup 0 --no-cpu-baseline > $O/r02b_ncu.log 2>&1
cat $O/r02b_umma_probe.txt $O/r02b_gather_probe.txt
