"""2+ GPU check of the fused peer-memory data-parallel step (run under torchrun):
   1. the same 3 optimisation steps through the fused kernel and through NCCL all-reduce + multi-tensor Adam must give the
      same parameters on every rank;
   2. prints which peer-memory backend was used and whether a flag barrier ever timed out."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import yaml  # noqa: E402

from hm_vae_b200 import ddp, ops  # noqa: E402
from hm_vae_b200.trainer_motion_vae import Trainer  # noqa: E402

rank, world, local = ddp.init_from_env("nccl")
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
cfg_name = sys.argv[1] if len(sys.argv) > 1 else "len64_no_aug_hm_vae.yaml"      # or trajectory_model.yaml (BASELINE config 5)
hp = yaml.safe_load(open(os.path.join(os.path.dirname(__file__), "..", "configs", cfg_name)))
bs, T = 8, hp["train_seq_len"]
g = torch.Generator().manual_seed(100 + rank)
x6 = torch.randn(bs, T, 24, 6, generator=g).to(dev)
rot = ops.rot6d_to_rotmat(x6)
data = (torch.stack((rot[..., 0], rot[..., 1]), dim=-2).reshape(bs, T, -1).contiguous(), rot.reshape(bs, T, -1).contiguous())
if hp["model_name"] == "TrajectoryModel":
    data = (data[0], data[1], None, torch.randn(bs, T, 72, generator=g).to(dev), None, None, torch.randn(bs, T, 3, generator=g).to(dev))
out, all_losses, start, first_m = {}, {}, None, {}
for fused in (False, True):
    torch.manual_seed(0)
    tr = Trainer(dict(hp), device=dev, sync_losses=False, dp_fused=fused).to(dev)
    ddp.broadcast_parameters(tr.model)
    if start is None:
        start = {k: v.detach().clone() for k, v in tr.model.named_parameters()}
    torch.manual_seed(50 + rank)
    losses = [float(tr.gen_update(data, hp, 0)[0])]
    torch.cuda.synchronize()
    # after ONE step exp_avg = (1 - beta1) * (mean gradient + wd * p0): the gradient each path applied, before Adam's
    # normalisation can amplify anything (state_dict is collective on the fused path: every rank calls it)
    first_m[fused] = {i: st["exp_avg"].detach().clone() for i, st in tr.gen_opt.state_dict()["state"].items()}
    losses += [float(tr.gen_update(data, hp, 0)[0]) for _ in range(2)]
    torch.cuda.synchronize()
    all_losses[fused] = losses
    out[fused] = {k: v.detach().clone() for k, v in tr.model.named_parameters()}
    if fused:
        print("rank %d: dp_mode %s timed_out %s losses %s" % (rank, tr.dp_mode, tr.gen_opt.timed_out(), losses), flush=True)
    else:
        print("rank %d: dp_mode %s losses %s" % (rank, tr.dp_mode, losses), flush=True)
    ops.unregister_grad_buffers()
# TIGHT part: the gradient applied at step 1 (reduce-scatter inside the fused kernel vs NCCL all-reduce), per tensor, 1e-5
# relative-L2 (only the cross-rank summation order differs: ~1e-7).
grad_rel = 0.0
assert set(first_m[True]) == set(first_m[False]) and first_m[True], "the two paths stepped different parameter sets"
for i in first_m[True]:
    a, b = first_m[True][i].double(), first_m[False][i].double().to(first_m[True][i].device)
    grad_rel = max(grad_rel, float((a - b).norm() / b.norm().clamp_min(1e-30)))
# What else is compared, and why the parameter bound after 3 steps is loose (the kernel-level proof is tools/dp_kernel_check.py):
# the two runs use two different Adam kernels (dp_adam_kernel vs the multi-tensor adam_kernel) whose results differ in the last
# bit (FMA contraction), so from step 2 on their losses / gradients differ at the 1e-7 level.  Adam's early updates are ~ +-lr
# whatever the gradient's magnitude, so the ~1 % of elements whose gradient is itself rounding noise (|g| ~ 1e-9, e.g. weights
# of taps that only ever see the reflect padding) can move in opposite directions: |dp| <= 2 * lr * steps for those.  Losses must
# agree (bounds below), every rank must hold bit-identical parameters, and the update as a whole must agree to 5 % relative-L2.
worst, differing, total = 0.0, 0, 0
num = den = 0.0
for k in out[False]:
    a, b = out[True][k].double(), out[False][k].double()
    d = (a - b).abs()
    num += float(((a - b) ** 2).sum())
    den += float(((b - start[k].double()) ** 2).sum())
    worst = max(worst, float(d.max()))
    differing += int((d > 1e-6).sum())
    total += d.numel()
# every rank must hold identical parameters after the fused steps
chk = torch.stack([v.double().sum() for v in out[True].values()]).sum().reshape(1)
gathered = [torch.zeros_like(chk) for _ in range(world)]
dist.all_gather(gathered, chk)
same = all(float(t) == float(gathered[0]) for t in gathered)
upd_rel = (num / max(den, 1e-300)) ** 0.5
loss_rel = max(abs(a - b) / abs(b) for a, b in zip(all_losses[True], all_losses[False]))
# The share of elements that differ by > 1e-6 after 3 steps is REPORTED, not bounded: it measures how many noise-level gradients
# the reduction order flipped at step 1 (+-lr each) and what the network made of that in steps 2-3 -- 1.3 % at 2 ranks (a + b is
# order-free), 14 % at 8 ranks (NCCL's ring vs the switch's tree), with the step-1 gradients equal to 1e-7 in both cases.
# Losses: step 1 runs on identical parameters (1e-6); steps 2-3 run on parameters that already carry the flipped +-lr updates,
# so they are held to 1e-4 (measured 1.2e-7 / 1.3e-5 / 1.1e-7 at 2 / 4 / 8 ranks; north_star's loss tolerance is 2e-3).
loss1_rel = abs(all_losses[True][0] - all_losses[False][0]) / abs(all_losses[False][0])
ok = grad_rel <= 1e-5 and loss1_rel <= 1e-6 and loss_rel <= 1e-4 and upd_rel <= 0.05 and worst <= 6.5e-4 and same
print("rank %d: fused vs nccl: step-1 applied gradient agrees to %.2e relative-L2 (worst tensor), step-1 loss to %.2e; losses of "
      "steps 1-3 agree to %.2e; update after 3 steps: relative-L2 difference %.3e, worst abs parameter diff %.3e (bound 2*lr*steps "
      "= 6e-4), %.4f%% of elements differ by > 1e-6; ranks identical: %s -- %s" % (
          rank, grad_rel, loss1_rel, loss_rel, upd_rel, worst, 100.0 * differing / total, same, "PASS" if ok else "FAIL"), flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
