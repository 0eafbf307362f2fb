set -x
mkdir -p gpurun_out
if [ "$1" != "notests" ]; then
python -m pytest tests -m gpu -x -q > gpurun_out/r1_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r1_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1_smoke.log 2>&1
fi
python bench.py > gpurun_out/r1_bench.json 2> gpurun_out/r1_bench.err; echo "bench rc=$?" >> gpurun_out/r1_bench.err
python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/r1_bench_ref.json 2>> gpurun_out/r1_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 390 -c 300 --csv --log-file gpurun_out/r1_launches.csv python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-fk-sweep --no-large-batch > gpurun_out/r1_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"conv_wgrad_tc_kernel|conv_tc_kernel|conv_pack_kernel|recon_kernel" -s 99 -c 34 -o /tmp/r1_conv_full -f python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-fk-sweep --no-large-batch > gpurun_out/r1_ncu_full.log 2>&1
ncu -i /tmp/r1_conv_full.ncu-rep --page raw --csv > gpurun_out/r1_conv_full_raw.csv 2>/dev/null
ncu -i /tmp/r1_conv_full.ncu-rep --page source --csv --print-source sass,cuda -k regex:conv_wgrad_tc_kernel -c 1 > gpurun_out/r1_wgrad_source.csv 2>/dev/null
ncu -i /tmp/r1_conv_full.ncu-rep --page source --csv -k regex:conv_tc_kernel -c 1 > gpurun_out/r1_convtc_source.csv 2>/dev/null
ncu -i /tmp/r1_conv_full.ncu-rep --page details -k regex:conv_wgrad_tc_kernel -c 1 > gpurun_out/r1_wgrad_details.txt 2>/dev/null
ncu -i /tmp/r1_conv_full.ncu-rep --page details -k regex:conv_tc_kernel -c 1 > gpurun_out/r1_convtc_details.txt 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:"fk_|rot6d" -s 4 -c 4 -o /tmp/r1_fk_full -f python tools/fk_probe.py > gpurun_out/r1_ncu_fk.log 2>&1
ncu -i /tmp/r1_fk_full.ncu-rep --page raw --csv > gpurun_out/r1_fk_full_raw.csv 2>/dev/null
ncu -i /tmp/r1_fk_full.ncu-rep --page details > gpurun_out/r1_fk_details.txt 2>/dev/null
ls -la /tmp/*.ncu-rep
du -sh gpurun_out
