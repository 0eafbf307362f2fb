#!/bin/bash
# usage: tools/dp_split_ab.sh N      multi-GPU bench under different split settings of the fused data-parallel optimiser
N=${1:-2}
run() {
  echo "== $*"
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) bench.py --gpus $N --steps 200 --warmup 10 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('  seq/s %.0f  ms/step %.4f  e2e %.0f  mode %s timeout %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['run']['data_parallel'], d['run'].get('dp_barrier_timed_out')))
"
}
run HMVAE_DP_SPLIT=0
run HMVAE_DP_SPLIT=1 HMVAE_DP_PARTIAL_CTAS=32
run HMVAE_DP_SPLIT=1 HMVAE_DP_PARTIAL_CTAS=74
run HMVAE_DP_SPLIT=1 HMVAE_DP_PARTIAL_CTAS=148
