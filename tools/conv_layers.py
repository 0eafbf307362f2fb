"""Runs every len64 conv layer (fprop + dgrad + wgrad through the module surface) a few times at a given batch, so that an
`ncu --metrics gpu__time_duration.sum --cache-control none --clock-control none` launch list gives WARM per-kernel device times
per layer (tools/ncu_layers.py summarises it).  Usage: python tools/conv_layers.py [batch] [reps]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import hm_vae_b200 as H  # noqa: E402
from hm_vae_b200 import ops  # noqa: E402

LAYERS = [  # (level, ci, co, K, stride, T_in, upsample, unpool_level)
    (0, 6, 12, 15, 2, 64, False, None), (1, 12, 24, 15, 2, 32, False, None), (2, 24, 48, 15, 2, 16, False, None),
    (3, 48, 96, 15, 2, 8, False, None), (3, 96, 48, 15, 1, 8, True, 3), (2, 48, 24, 15, 1, 16, True, 2),
    (1, 24, 12, 15, 1, 32, True, 1), (0, 24, 6, 15, 1, 64, True, 0)]

b = int(sys.argv[1]) if len(sys.argv) > 1 else 32
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
topo = json.load(open(os.path.join(ROOT, "tests", "golden", "topology.json")))["levels"]
dev = torch.device("cuda", 0)
torch.manual_seed(0)
for li, (lvl, ci, co, k, s, t_in, up, unpool_lvl) in enumerate(LAYERS):
    nb = topo[lvl]["neighbours"]
    j = len(nb)
    conv = H.SkeletonConv(nb, j * ci, j * co, k, j, stride=s, padding=(k - 1) // 2, bias=True, padding_mode="reflection").to(dev)
    if unpool_lvl is not None:
        pl = topo[unpool_lvl]["pooling_list"]
        un = H.SkeletonUnpool(pl, ci)
        x = torch.randn(b, len(pl) * ci, t_in // 2, device=dev, requires_grad=True)
        kw = dict(upsample=True, unpool_src=un.src, src_joints=len(pl), lrelu=True)
    else:
        x = torch.randn(b, j * ci, t_in, device=dev, requires_grad=True)
        kw = dict(lrelu=True)
    gy = None
    for _ in range(reps):
        y = conv.fused_forward(x, **kw)
        if gy is None:
            gy = torch.randn_like(y)
        y.backward(gy)
        x.grad = None
        conv.weight.grad = None
        conv.bias.grad = None
    torch.cuda.synchronize()
    print("layer", li, "ok", tuple(y.shape), flush=True)
