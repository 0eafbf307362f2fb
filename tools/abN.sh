# multi-GPU A/B: bash tools/abN.sh N [check]
N=$1
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
B="bench.py --gpus $N --steps 200 --warmup 5 --no-cpu-baseline --no-fk-sweep --no-large-batch"
if [ "$2" = "check" ]; then timeout 200 $T --master-port 29590 tools/dp_check.py > gpurun_out/abN_check.log 2>&1; echo "rc=$?" >> gpurun_out/abN_check.log; fi
HMVAE_DP_SPLIT=1 timeout 200 $T --master-port 29591 $B > gpurun_out/abN_split.json 2> gpurun_out/abN_split.err
HMVAE_DP_SPLIT=0 timeout 200 $T --master-port 29592 $B > gpurun_out/abN_nosplit.json 2> gpurun_out/abN_nosplit.err
