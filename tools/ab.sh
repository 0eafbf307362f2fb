# A/B helper: runs bench.py several times inside ONE gpurun call (box-to-box variation is ~3 %, run-to-run on one box < 0.5 % at
# 200 steps).  Edit the variable assignments; results land in gpurun_out/ab_*.json (tools/show_bench.py prints them).
python -m pytest tests -m gpu -x -q > gpurun_out/ab_tests.log 2>&1; echo "rc=$?" >> gpurun_out/ab_tests.log
B="python bench.py --steps 200 --warmup 5 --no-cpu-baseline --no-fk-sweep --no-large-batch"
HMVAE_DP_SPLIT=0 $B > gpurun_out/ab_a.json 2> gpurun_out/ab_a.err
HMVAE_DP_SPLIT=1 $B > gpurun_out/ab_b.json 2> gpurun_out/ab_b.err
HMVAE_DP_SPLIT=1 HMVAE_DP_PARTIAL_CTAS=74 $B > gpurun_out/ab_c.json 2> gpurun_out/ab_c.err
HMVAE_DP_SPLIT=1 HMVAE_DP_PARTIAL_CTAS=296 $B > gpurun_out/ab_d.json 2> gpurun_out/ab_d.err
