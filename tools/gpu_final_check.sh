python -m pytest tests -m gpu -x -q > gpurun_out/f_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/f_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/f_smoke.log
python bench.py --steps 200 --warmup 10 > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err; echo "bench rc=$?" >> gpurun_out/f_bench.err
python bench.py > gpurun_out/f_bench_default.json 2> gpurun_out/f_bench_default.err; echo "bench rc=$?" >> gpurun_out/f_bench_default.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 390 -c 240 --csv --log-file gpurun_out/f_launches.csv python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-fk-sweep --no-large-batch > gpurun_out/f_ncu_launch.log 2>&1
