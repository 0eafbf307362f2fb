"""Two eager training steps (len64, B=32) between cudaProfilerStart/Stop, for
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/step_profile.py
Without ncu it just runs (and prints the step's launch count).  HMVAE_STACK=0 selects the per-layer path."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import yaml  # noqa: E402

from hm_vae_b200 import _lib, ops, stack  # noqa: E402
from hm_vae_b200.trainer_motion_vae import Trainer  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
stack.set_enabled(os.environ.get("HMVAE_STACK", "1") != "0")
hp = yaml.safe_load(open(os.path.join(ROOT, "configs", "len64_no_aug_hm_vae.yaml")))
bs, T = int(os.environ.get("BATCH", "32")), hp["train_seq_len"]
torch.manual_seed(0)
tr = Trainer(dict(hp), device=dev, sync_losses=False).to(dev)
g = torch.Generator().manual_seed(1234)
rot = ops.rot6d_to_rotmat(torch.randn(bs, T, 24, 6, generator=g).to(dev))
data = (torch.stack((rot[..., 0], rot[..., 1]), dim=-2).reshape(bs, T, -1).contiguous(), rot.reshape(bs, T, -1).contiguous())
for _ in range(5):
    tr.gen_update(data, hp, 0)
torch.cuda.synchronize()
n0 = _lib.launch_count()
torch.cuda.profiler.start()
for _ in range(2):
    tr.gen_update(data, hp, 0)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("hmvae launches per step:", (_lib.launch_count() - n0) // 2, "dp_mode", tr.dp_mode)
