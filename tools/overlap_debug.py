import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, yaml
from hm_vae_b200 import ops
from hm_vae_b200.seq_two_hier_sa_vae import TwoHierSAVAEModel
hp = yaml.safe_load(open(os.path.join(ROOT, "configs", "len64_no_aug_hm_vae.yaml")))
torch.manual_seed(0)
model = TwoHierSAVAEModel(dict(hp), device="cuda").cuda()
g = torch.Generator().manual_seed(1)
x6 = torch.randn(4, 64, 24, 6, generator=g).cuda()
rot = ops.rot6d_to_rotmat(x6)
d6 = torch.stack((rot[..., 0], rot[..., 1]), dim=-2).reshape(4, 64, -1).contiguous()
dm = rot.reshape(4, 64, -1).contiguous()
eps = [torch.randn(4 * k, d, generator=g).cuda() for k, d in [(14, 12), (9, 24), (7, 24), (7, 24)]]
def run(overlap, sync_before_read, impl):
    ops.set_conv_impl(impl)
    model.zero_grad(set_to_none=True)
    if not overlap:
        orig = ops.wgrad_overlap.__enter__
        ops.wgrad_overlap.__enter__ = lambda self: self
    try:
        model((d6, dm), hp, 0, eps_list=eps)
    finally:
        if not overlap:
            ops.wgrad_overlap.__enter__ = orig
    if sync_before_read:
        torch.cuda.synchronize()
    return {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None and not n.startswith("dec.enc")}
for impl in (1, 0):
    ref = run(False, True, impl)
    for ov, sy in [(True, True), (True, False)]:
        got = run(ov, sy, impl)
        torch.cuda.synchronize()
        worst = max(((got[k] - ref[k]).norm() / (ref[k].norm() + 1e-30)).item() for k in ref)
        bad = [k for k in ref if ((got[k] - ref[k]).norm() / (ref[k].norm() + 1e-30)).item() > 1e-3]
        print("impl", impl, "overlap", ov, "sync", sy, "worst rel", worst, bad[:4])
