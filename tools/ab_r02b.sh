# A/B of this session's step-scheduling changes inside ONE gpurun call (1 GPU): mask-aware optimiser table, second optimiser
# bucket (deepest encoder level), N(0,1) draws issued after the encoder kernels.  Results: gpurun_out/ab2_*.json
B="python bench.py --steps 200 --warmup 5 --no-cpu-baseline --no-fk-sweep --no-large-batch --no-reference-cuda --no-other-configs"
python -m pytest tests -m gpu -x -q > gpurun_out/ab2_tests.log 2>&1; echo "rc=$?" >> gpurun_out/ab2_tests.log
HMVAE_DP_MASK_AWARE=0 HMVAE_DP_SPLIT_ENC=0 HMVAE_LATE_EPS=0 $B > gpurun_out/ab2_a_old.json 2> gpurun_out/ab2_a.err
HMVAE_DP_SPLIT_ENC=0 HMVAE_LATE_EPS=0 $B > gpurun_out/ab2_b_mask.json 2> gpurun_out/ab2_b.err
HMVAE_LATE_EPS=0 $B > gpurun_out/ab2_c_mask_bucket2.json 2> gpurun_out/ab2_c.err
$B > gpurun_out/ab2_d_all.json 2> gpurun_out/ab2_d.err
HMVAE_DP_PARTIAL_CTAS=148 $B > gpurun_out/ab2_e_all_148.json 2> gpurun_out/ab2_e.err
HMVAE_DP_IN_FLIGHT_LOCAL=2 $B > gpurun_out/ab2_f_all_u2.json 2> gpurun_out/ab2_f.err
python tools/timeline.py > gpurun_out/ab2_timeline.txt 2> gpurun_out/ab2_timeline.err
