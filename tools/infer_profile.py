"""BASELINE config 4 (len64 `test()` path, B=512, no grad) between cudaProfilerStart/Stop, for ncu --profile-from-start off."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import yaml  # noqa: E402

from hm_vae_b200 import ops  # noqa: E402
from hm_vae_b200.seq_two_hier_sa_vae import TwoHierSAVAEModel  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
hp = yaml.safe_load(open(os.path.join(ROOT, "configs", "len_64_test_interpolation.yaml")))
bs, T = int(os.environ.get("BATCH", "512")), hp["train_seq_len"]
torch.manual_seed(0)
model = TwoHierSAVAEModel(dict(hp), device=dev).to(dev)
g = torch.Generator().manual_seed(99)
rot = ops.rot6d_to_rotmat(torch.randn(bs, T, 24, 6, generator=g).to(dev))
data = (torch.stack((rot[..., 0], rot[..., 1]), dim=-2).reshape(bs, T, -1).contiguous(), rot.reshape(bs, T, -1).contiguous())
for _ in range(3):
    model.test(data, hp, 0)
torch.cuda.synchronize()
torch.cuda.profiler.start()
model.test(data, hp, 0)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok")
