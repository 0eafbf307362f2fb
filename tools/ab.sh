B="python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-fk-sweep --no-large-batch"
HMVAE_TC_STAGES=2 HMVAE_WG_STAGES=2 HMVAE_WG_TMEM_COLS=256 $B > gpurun_out/ab_a.json 2> gpurun_out/ab_a.err
HMVAE_TC_STAGES=2 HMVAE_WG_STAGES=2 HMVAE_WG_TMEM_COLS=128 $B > gpurun_out/ab_b.json 2> gpurun_out/ab_b.err
HMVAE_TC_STAGES=2 HMVAE_WG_STAGES=3 HMVAE_WG_TMEM_COLS=256 $B > gpurun_out/ab_c.json 2> gpurun_out/ab_c.err
HMVAE_TC_STAGES=2 HMVAE_WG_STAGES=4 HMVAE_WG_TMEM_COLS=256 $B > gpurun_out/ab_d.json 2> gpurun_out/ab_d.err
HMVAE_TC_STAGES=2 HMVAE_WG_STAGES=2 HMVAE_WG_TMEM_COLS=384 $B > gpurun_out/ab_e.json 2> gpurun_out/ab_e.err
