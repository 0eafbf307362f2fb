#!/bin/bash
# usage: tools/gpu_multi_r02b.sh N : Trainer-level parity (fused step vs NCCL all-reduce) and the bucketed-overlap A/B at N ranks
N=${1:-8}
tr() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) "${@:2}"; }
show() { python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('  seq/s %.0f  ms/step %.4f  e2e %.0f  mode %s timeout %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['run']['data_parallel'], d['run'].get('dp_barrier_timed_out')))
"; }
{
  echo "=== dp_check (Trainer, 3 steps, fused vs NCCL all-reduce) W=$N default path"; tr $N tools/dp_check.py 2>&1 | grep -E "^rank 0|Error|error|Traceback" | head -20; echo "rc=${PIPESTATUS[0]}"
  echo "=== dp_check W=$N HMVAE_DP_SPLIT=1 (bucketed overlap)"; HMVAE_DP_SPLIT=1 HMVAE_DP_PARTIAL_CTAS=32 HMVAE_DP_PARTIAL_IN_FLIGHT=4 tr $N tools/dp_check.py 2>&1 | grep -E "^rank 0|Error|error|Traceback" | head -20; echo "rc=${PIPESTATUS[0]}"
} > gpurun_out/m2_dp_check_${N}gpu.log 2>&1
{
  B="bench.py --gpus $N --steps 150 --warmup 10"
  echo "== split off"; tr $N $B 2>/dev/null | show
  echo "== HMVAE_DP_SPLIT=1 CTAS=32 IN_FLIGHT=4 (two buckets)"; HMVAE_DP_SPLIT=1 HMVAE_DP_PARTIAL_CTAS=32 HMVAE_DP_PARTIAL_IN_FLIGHT=4 tr $N $B 2>/dev/null | show
  echo "== HMVAE_DP_SPLIT=1 CTAS=64 IN_FLIGHT=2 (two buckets)"; HMVAE_DP_SPLIT=1 HMVAE_DP_PARTIAL_CTAS=64 HMVAE_DP_PARTIAL_IN_FLIGHT=2 tr $N $B 2>/dev/null | show
  echo "== HMVAE_DP_SPLIT=1 CTAS=32 IN_FLIGHT=4, decoder bucket only"; HMVAE_DP_SPLIT=1 HMVAE_DP_SPLIT_ENC=0 HMVAE_DP_PARTIAL_CTAS=32 HMVAE_DP_PARTIAL_IN_FLIGHT=4 tr $N $B 2>/dev/null | show
} > gpurun_out/m2_overlap_ab_${N}gpu.log 2>&1
