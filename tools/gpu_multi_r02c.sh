#!/bin/bash
# usage: tools/gpu_multi_r02c.sh N : Trainer-level parity with the step-1 gradient bound at N and 2 ranks, then the decoder-bucket
# overlap on / off at N ranks (default policy: on when the NVSwitch multicast path is in use)
N=${1:-4}
tr() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) "${@:2}"; }
show() { python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('  seq/s %.0f  ms/step %.4f  e2e %.0f  mode %s timeout %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['run']['data_parallel'], d['run'].get('dp_barrier_timed_out')))
"; }
{
  for W in $N 2; do
    echo "=== dp_check (Trainer, 3 steps, fused vs NCCL all-reduce) W=$W default policy"; tr $W tools/dp_check.py > /tmp/dpc.log 2>&1; rc=$?
    grep -E "^rank 0|Traceback|Error" /tmp/dpc.log | head -8; echo "rc=$rc"
  done
} > gpurun_out/m3_dp_check_${N}gpu.log 2>&1
{
  B="bench.py --gpus $N --steps 150 --warmup 10"
  echo "== default (auto)"; tr $N $B 2>/dev/null | show
  echo "== HMVAE_DP_SPLIT=0"; HMVAE_DP_SPLIT=0 tr $N $B 2>/dev/null | show
} > gpurun_out/m3_overlap_ab_${N}gpu.log 2>&1
