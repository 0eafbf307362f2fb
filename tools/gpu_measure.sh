# Round measurement pass on one B200 (see profiles/r01_summary_v3.md).  Usage: bash tools/gpu_measure.sh [notests]
set -x
mkdir -p gpurun_out
if [ "$1" != "notests" ]; then
python -m pytest tests -m gpu -x -q > gpurun_out/m_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/m_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/m_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/m_smoke.log
fi
python bench.py --steps 200 --warmup 10 > gpurun_out/m_bench.json 2> gpurun_out/m_bench.err; echo "bench rc=$?" >> gpurun_out/m_bench.err
python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/m_bench_ref.json 2>> gpurun_out/m_bench.err
NCUB="python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-fk-sweep --no-large-batch"
ncu --metrics gpu__time_duration.sum --clock-control none -s 390 -c 240 --csv --log-file gpurun_out/m_launches.csv $NCUB > gpurun_out/m_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"conv_wgrad_tc_kernel|conv_tc_kernel|conv_pack_kernel|recon_par_kernel|dp_adam_kernel" -s 99 -c 36 -o /tmp/m_conv_full -f $NCUB > gpurun_out/m_ncu_full.log 2>&1
ncu -i /tmp/m_conv_full.ncu-rep --page raw --csv > gpurun_out/m_conv_full_raw.csv 2>/dev/null
ncu -i /tmp/m_conv_full.ncu-rep --page details -k regex:conv_wgrad_tc_kernel -c 1 > gpurun_out/m_wgrad_details.txt 2>/dev/null
ncu -i /tmp/m_conv_full.ncu-rep --page details -k regex:conv_tc_kernel -c 1 > gpurun_out/m_convtc_details.txt 2>/dev/null
ncu -i /tmp/m_conv_full.ncu-rep --page details -k regex:dp_adam_kernel -c 1 > gpurun_out/m_dpadam_details.txt 2>/dev/null
ncu -i /tmp/m_conv_full.ncu-rep --page details -k regex:recon_par_kernel -c 1 > gpurun_out/m_recon_details.txt 2>/dev/null
python tools/fk_bench.py > gpurun_out/m_fk.json 2> gpurun_out/m_fk.err
du -sh gpurun_out
