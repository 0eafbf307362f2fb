#!/bin/bash
# B=512 test() path (BASELINE config 4) under different conv_tc tiling knobs; prints conv TFLOP/s and the per-layer device times
run() {
  echo "== $*"
  env "$@" timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-fk-sweep --no-reference-cuda --no-other-configs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readlines()[-1]); c=d['conv_large_batch']
print('  ms/call %.3f conv_ms %.3f TF/s %.1f frac %.3f' % (c['ms_per_call'], c['conv_fprop_ms'], c['conv_tflops'], c['conv_frac_of_tf32_peak']))
L=c['conv_us_per_layer']
print('  run  ', [L.get('fprop_tc_run[L%d]'%i) for i in range(8)])
print('  link ', [L.get('fprop_link[L%d]'%i) for i in range(8)])
print('  step ms (B=32): %.4f' % d['ms_per_step'])
"
}
run A=1
run HMVAE_TC_GROUP_COLS=128 HMVAE_TC_GROUP_MAX=8
run HMVAE_TC_GROUP_COLS=128 HMVAE_TC_GROUP_MAX=8 HMVAE_TC_STAGES=3
run HMVAE_TC_GROUP_COLS=96 HMVAE_TC_GROUP_MAX=6
run HMVAE_TC_STAGES=3
run HMVAE_TC_STAGES=4
