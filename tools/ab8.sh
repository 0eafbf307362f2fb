N=$1
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
B="bench.py --gpus $N --steps 200 --warmup 5 --no-cpu-baseline --no-fk-sweep --no-large-batch"
HMVAE_DP_MULTICAST=1 timeout 200 $T --master-port 29581 $B > gpurun_out/r20_g${N}_mc.json 2> gpurun_out/r20_g${N}_mc.err
if [ "$2" = "both" ]; then HMVAE_DP_MULTICAST=0 timeout 200 $T --master-port 29582 $B > gpurun_out/r20_g${N}_uc.json 2> gpurun_out/r20_g${N}_uc.err; fi
