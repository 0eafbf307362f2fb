T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 200 $T --master-port 29571 tools/dp_check.py > gpurun_out/r18_dpcheck2.log 2>&1; echo "rc=$?" >> gpurun_out/r18_dpcheck2.log
B="bench.py --gpus 2 --steps 200 --warmup 5 --no-cpu-baseline --no-fk-sweep --no-large-batch"
timeout 200 $T --master-port 29572 $B > gpurun_out/r18_g2_mc.json 2> gpurun_out/r18_g2_mc.err
HMVAE_DP_MULTICAST=0 timeout 200 $T --master-port 29573 $B > gpurun_out/r18_g2_uc.json 2> gpurun_out/r18_g2_uc.err
