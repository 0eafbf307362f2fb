B="python bench.py --steps 200 --warmup 5 --no-cpu-baseline --no-fk-sweep --no-large-batch"
HMVAE_PDL=0 $B > gpurun_out/ab_a.json 2> gpurun_out/ab_a.err
HMVAE_PDL=2 $B > gpurun_out/ab_b.json 2> gpurun_out/ab_b.err
HMVAE_PDL=1 $B > gpurun_out/ab_c.json 2> gpurun_out/ab_c.err
HMVAE_PDL=0 $B > gpurun_out/ab_d.json 2> gpurun_out/ab_d.err
HMVAE_PDL=2 $B > gpurun_out/ab_e.json 2> gpurun_out/ab_e.err
