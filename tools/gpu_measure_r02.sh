# Round-2 ncu pass on one B200 (summaries go to profiles/r02_*).  Every profiled command first exits 0 without ncu.
set -x
mkdir -p gpurun_out
python tools/step_profile.py > gpurun_out/q_step_plain.log 2>&1 &&
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/q_launches_step.csv python tools/step_profile.py > gpurun_out/q_ncu_launch.log 2>&1
python tools/infer_profile.py > gpurun_out/q_infer_plain.log 2>&1 &&
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/q_launches_infer512.csv python tools/infer_profile.py > gpurun_out/q_ncu_launch2.log 2>&1
# full captures: training step (B=32) -- one instance of each hot kernel class
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"conv_tc_kernel|conv_link_kernel|conv_wgrad_tc_kernel|heads_fwd_kernel|heads_bwd_kernel|dp_adam_kernel|recon_par_kernel" -c 60 -o /tmp/q_step_full -f python tools/step_profile.py > gpurun_out/q_ncu_full.log 2>&1
ncu -i /tmp/q_step_full.ncu-rep --page raw --csv > gpurun_out/q_step_full_raw.csv 2>/dev/null
for k in conv_tc_kernel conv_link_kernel conv_wgrad_tc_kernel heads_fwd_kernel dp_adam_kernel; do
  ncu -i /tmp/q_step_full.ncu-rep --page details -k regex:$k -c 1 > gpurun_out/q_details_$k.txt 2>/dev/null
done
# full capture: B=512 inference convs (tensor-pipe utilisation evidence)
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"conv_tc_kernel|conv_link_kernel" -c 24 -o /tmp/q_infer_full -f python tools/infer_profile.py > gpurun_out/q_ncu_full2.log 2>&1
ncu -i /tmp/q_infer_full.ncu-rep --page raw --csv > gpurun_out/q_infer512_full_raw.csv 2>/dev/null
ncu -i /tmp/q_infer_full.ncu-rep --page details -k regex:conv_tc_kernel -c 1 > gpurun_out/q_details_infer512_conv_tc.txt 2>/dev/null
# FK kernels
python tools/fk_bench.py > gpurun_out/q_fk.json 2> gpurun_out/q_fk.err
du -sh gpurun_out
