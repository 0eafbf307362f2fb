#!/bin/bash
# usage: tools/dp_overlap_ab.sh N     A/B of the fused data-parallel optimiser on N GPUs inside ONE gpurun call:
# mask-aware unit table on / off, decoder share issued under the encoder backward (split) with few CTAs and deep loads in flight
N=${1:-2}
run() {
  echo "== $*"
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) bench.py --gpus $N --steps 200 --warmup 10 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('  seq/s %.0f  ms/step %.4f  e2e %.0f  mode %s timeout %s skipped %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['run']['data_parallel'], d['run'].get('dp_barrier_timed_out'), d['run'].get('optimiser_masked_elems_skipped')))
"
}
run HMVAE_DP_SPLIT=0 HMVAE_DP_MASK_AWARE=0
run HMVAE_DP_SPLIT=0
run HMVAE_DP_SPLIT=1 HMVAE_DP_PARTIAL_CTAS=64 HMVAE_DP_PARTIAL_IN_FLIGHT=8
run HMVAE_DP_SPLIT=1 HMVAE_DP_PARTIAL_CTAS=32 HMVAE_DP_PARTIAL_IN_FLIGHT=8
run HMVAE_DP_SPLIT=1 HMVAE_DP_PARTIAL_CTAS=96 HMVAE_DP_PARTIAL_IN_FLIGHT=4
run HMVAE_DP_SPLIT=0 HMVAE_DP_IN_FLIGHT=4
run HMVAE_DP_SPLIT=0 HMVAE_DP_MULTICAST=1
run HMVAE_DP_SPLIT=1 HMVAE_DP_MULTICAST=1 HMVAE_DP_PARTIAL_CTAS=64 HMVAE_DP_PARTIAL_IN_FLIGHT=8
