# A/B inside ONE gpurun call (1 GPU): issue order of the weight gradients, last weight gradient on the chain's stream, staging-pass
# residency.  Results: gpurun_out/ab3_*.json
B="python bench.py --steps 200 --warmup 5 --no-cpu-baseline --no-fk-sweep --no-large-batch --no-reference-cuda --no-other-configs"
python -m pytest tests -m gpu -x -q > gpurun_out/ab3_tests.log 2>&1; echo "rc=$?" >> gpurun_out/ab3_tests.log
HMVAE_WGRAD_LAST_INLINE=0 HMVAE_WGRAD_AFTER_DGRAD=0 $B > gpurun_out/ab3_a_base.json 2> gpurun_out/ab3_a.err
HMVAE_WGRAD_LAST_INLINE=1 HMVAE_WGRAD_AFTER_DGRAD=0 $B > gpurun_out/ab3_b_inline.json 2> gpurun_out/ab3_b.err
HMVAE_WGRAD_LAST_INLINE=0 HMVAE_WGRAD_AFTER_DGRAD=1 $B > gpurun_out/ab3_c_after.json 2> gpurun_out/ab3_c.err
$B > gpurun_out/ab3_d_both.json 2> gpurun_out/ab3_d.err
HMVAE_WG_PREP_CTAS_PER_SM=4 $B > gpurun_out/ab3_e_both_prep4.json 2> gpurun_out/ab3_e.err
HMVAE_WG_PREP_CTAS_PER_SM=2 $B > gpurun_out/ab3_f_both_prep2.json 2> gpurun_out/ab3_f.err
HMVAE_WGRAD_STREAMS=3 $B > gpurun_out/ab3_g_both_3streams.json 2> gpurun_out/ab3_g.err
HMVAE_DP_SPLIT_ENC=0 $B > gpurun_out/ab3_h_both_nobucket2.json 2> gpurun_out/ab3_h.err
python tools/timeline.py > gpurun_out/ab3_timeline.txt 2> gpurun_out/ab3_timeline.err
