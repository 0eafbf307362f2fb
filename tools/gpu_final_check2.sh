# final confirmation (1 GPU): GPU tests with and without the no_grad boundary-tensor skip, smoke, default bench line, launch list
python -m pytest tests -m gpu -x -q > gpurun_out/f_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/f_tests.log
HMVAE_STACK_SKIP_BOUNDS=0 python -m pytest tests -m gpu -x -q -k "b512 or stack_path or test_path" > gpurun_out/f_tests_noskip.log 2>&1; echo "tests rc=$?" >> gpurun_out/f_tests_noskip.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/f_smoke.log
python bench.py > gpurun_out/f_bench_default.json 2> gpurun_out/f_bench_default.err; echo "bench rc=$?" >> gpurun_out/f_bench_default.err
HMVAE_STACK_SKIP_BOUNDS=0 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-fk-sweep --no-reference-cuda --no-other-configs > gpurun_out/f_bench_noskip.json 2> gpurun_out/f_bench_noskip.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 390 -c 270 --csv --log-file gpurun_out/f_launches.csv python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-fk-sweep --no-large-batch --no-reference-cuda --no-other-configs > gpurun_out/f_ncu_launch.log 2>&1
