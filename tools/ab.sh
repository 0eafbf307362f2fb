python -m pytest tests -m gpu -x -q > gpurun_out/r12_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r12_tests.log
B="python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-fk-sweep --no-large-batch"
$B > gpurun_out/ab_a.json 2> gpurun_out/ab_a.err
$B > gpurun_out/ab_b.json 2> gpurun_out/ab_b.err
