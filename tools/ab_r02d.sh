# A/B inside ONE gpurun call (1 GPU): weight-gradient x tiles staged during the forward pass.  Results: gpurun_out/ab4_*.json
B="python bench.py --steps 200 --warmup 5 --no-cpu-baseline --no-fk-sweep --no-large-batch --no-reference-cuda --no-other-configs"
python -m pytest tests -m gpu -x -q > gpurun_out/ab4_tests.log 2>&1; echo "rc=$?" >> gpurun_out/ab4_tests.log
HMVAE_WGRAD_PRESTAGE=0 $B > gpurun_out/ab4_a_off.json 2> gpurun_out/ab4_a.err
$B > gpurun_out/ab4_b_on.json 2> gpurun_out/ab4_b.err
HMVAE_WGRAD_PRESTAGE=0 $B > gpurun_out/ab4_c_off.json 2> gpurun_out/ab4_c.err
$B > gpurun_out/ab4_d_on.json 2> gpurun_out/ab4_d.err
HMVAE_WG_PREP_CTAS_PER_SM=4 $B > gpurun_out/ab4_e_on_prep4.json 2> gpurun_out/ab4_e.err
python tools/timeline.py > gpurun_out/ab4_timeline.txt 2> gpurun_out/ab4_timeline.err
