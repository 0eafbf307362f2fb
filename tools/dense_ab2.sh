for np in 16 24 48; do
  echo "== HMVAE_TC_DENSE_MAXNP=$np"
  HMVAE_TC_DENSE_MAXNP=$np AB_ONLY=stack timeout 200 python tools/stack_ab.py 2>&1 | tail -1
  HMVAE_TC_DENSE_MAXNP=$np timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-fk-sweep --no-reference-cuda --no-other-configs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readlines()[-1]); c=d['conv_large_batch']
print('  B=512: ms/call %.3f conv_ms %.3f TF/s %.1f frac %.3f' % (c['ms_per_call'], c['conv_fprop_ms'], c['conv_tflops'], c['conv_frac_of_tf32_peak']))
L=c['conv_us_per_layer']
print('  run  ', [L.get('fprop_tc_run[L%d]'%i) for i in range(8)])
"
done
HMVAE_TC_DENSE_MAXNP=48 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "conv_layer or stack_path" 2>&1 | tail -2
