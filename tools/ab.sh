# A/B helper: runs bench.py several times inside ONE gpurun call (box-to-box variation is ~3 %, run-to-run on one box < 0.5 % at
# 200 steps).  Edit the variable assignments; results land in gpurun_out/ab_*.json (tools/show_bench.py prints them).
B="python bench.py --steps 200 --warmup 5 --no-cpu-baseline --no-fk-sweep --no-large-batch"
$B > gpurun_out/ab_a.json 2> gpurun_out/ab_a.err
HMVAE_TC_GROUP_COLS=32 HMVAE_TC_TARGET_CTAS=444 $B > gpurun_out/ab_b.json 2> gpurun_out/ab_b.err
HMVAE_TC_GROUP_COLS=32 HMVAE_TC_TARGET_CTAS=370 $B > gpurun_out/ab_c.json 2> gpurun_out/ab_c.err
HMVAE_TC_GROUP_COLS=32 HMVAE_TC_TARGET_CTAS=444 HMVAE_TC_MIN_SPLIT_LEN=2 $B > gpurun_out/ab_d.json 2> gpurun_out/ab_d.err
$B > gpurun_out/ab_e.json 2> gpurun_out/ab_e.err
