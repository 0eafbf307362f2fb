"""Runs the FK / rot6d kernels once at the 16 M joint-frame size (for `ncu --set full` captures of the HBM-bound kernels)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from hm_vae_b200._lib import check, int_array, lib, ptr, stream  # noqa: E402
from hm_vae_b200.fk_layer import load_smpl24  # noqa: E402

dev = torch.device("cuda", 0)
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 699051
parents, offsets, _ = load_smpl24()
par = int_array(parents)
off = torch.as_tensor(offsets, dtype=torch.float32, device=dev).contiguous()
x6 = torch.randn(frames, 24, 6, device=dev)
rot = torch.empty(frames, 24, 3, 3, device=dev)
gp = torch.randn(frames, 24, 3, device=dev)
gr = torch.randn(frames, 24, 3, 3, device=dev)
pos, grot, gx6 = torch.empty(frames, 24, 3, device=dev), torch.empty_like(rot), torch.empty_like(x6)
for _ in range(2):
    check(lib.hmvae_rot6d_fwd(ptr(x6), ptr(rot), frames * 24, stream()))
    check(lib.hmvae_fk_fwd(ptr(rot), 9, ptr(off), None, par, 24, frames, ptr(pos), None, stream()))
    check(lib.hmvae_fk_bwd(ptr(rot), 9, ptr(off), None, par, 24, frames, ptr(gp), None, ptr(grot), stream()))
    check(lib.hmvae_rot6d_bwd(ptr(x6), ptr(gr), ptr(gx6), frames * 24, stream()))
torch.cuda.synchronize()
print("ok", float(pos.abs().mean()))
