mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "conv or stack_path or golden or full_batch or b512" > gpurun_out/o_tests.log 2>&1; echo "rc=$?" >> gpurun_out/o_tests.log; tail -3 gpurun_out/o_tests.log
for d in 0 -1; do
  echo "== HMVAE_TC_DENSE=$d"
  HMVAE_TC_DENSE=$d AB_ONLY=stack timeout 200 python tools/stack_ab.py 2>&1 | tail -1
  HMVAE_TC_DENSE=$d timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-fk-sweep --no-reference-cuda --no-other-configs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readlines()[-1]); c=d['conv_large_batch']
print('  B=512: ms/call %.3f conv_ms %.3f TF/s %.1f frac %.3f' % (c['ms_per_call'], c['conv_fprop_ms'], c['conv_tflops'], c['conv_frac_of_tf32_peak']))
L=c['conv_us_per_layer']
print('  run  ', [L.get('fprop_tc_run[L%d]'%i) for i in range(8)])
print('  link ', [L.get('fprop_link[L%d]'%i) for i in range(8)])
"
done
HMVAE_TC_DENSE=-1 timeout 200 python tools/tc_phases.py 2>&1 | grep -v "enc0 dgrad"
