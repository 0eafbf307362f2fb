#!/bin/bash
# usage: tools/gpu_multi_r02.sh N   (N = 4 or 8 GPUs of one box): multi-rank parity of the fused optimiser step on the NVSwitch
# multicast path the scaling run uses (and on the unicast path), then the bench line at N and N/2 ranks.  Logs: gpurun_out/m_*
N=${1:-8}
tr() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) "${@:2}"; }
{
  for W in $N $((N / 2)); do
    echo "=== dp_kernel_check W=$W HMVAE_DP_MULTICAST=1 (NVLS)"; HMVAE_DP_MULTICAST=1 tr $W tools/dp_kernel_check.py 2>&1 | grep -E "rank|Error|error" ; echo "rc=${PIPESTATUS[0]}"
  done
  echo "=== dp_kernel_check W=$N HMVAE_DP_MULTICAST=0 (unicast)"; HMVAE_DP_MULTICAST=0 tr $N tools/dp_kernel_check.py 2>&1 | grep -E "rank|Error|error"; echo "rc=${PIPESTATUS[0]}"
  echo "=== dp_check (Trainer, 3 steps, fused vs NCCL all-reduce) W=$N default path"; tr $N tools/dp_check.py 2>&1 | tail -12; echo "rc=${PIPESTATUS[0]}"
} > gpurun_out/m_dp_parity_${N}gpu.log 2>&1
for W in $N $((N / 2)); do
  tr $W bench.py --gpus $W --steps 200 --warmup 10 > gpurun_out/m_bench_${W}gpu.json 2> gpurun_out/m_bench_${W}gpu.err
done
python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-fk-sweep --no-large-batch --no-reference-cuda --no-other-configs > gpurun_out/m_bench_1gpu.json 2> gpurun_out/m_bench_1gpu.err
