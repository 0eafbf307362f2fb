"""Device time (CUDA graph replay, no host gaps) of the conv stacks with the linked path on / off:
   encoder forward, decoder forward, and the whole training step, len64 at BATCH (default 32).  One JSON line."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import yaml  # noqa: E402

from hm_vae_b200 import ops, stack  # noqa: E402
from hm_vae_b200.trainer_motion_vae import Trainer  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
hp = yaml.safe_load(open(os.path.join(ROOT, "configs", "len64_no_aug_hm_vae.yaml")))
bs, T = int(os.environ.get("BATCH", "32")), hp["train_seq_len"]
g = torch.Generator().manual_seed(1234)
rot = ops.rot6d_to_rotmat(torch.randn(bs, T, 24, 6, generator=g).to(dev))
data = (torch.stack((rot[..., 0], rot[..., 1]), dim=-2).reshape(bs, T, -1).contiguous(), rot.reshape(bs, T, -1).contiguous())


def graph_time(fn, reps=20, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(reps):
            fn()
    gr.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        gr.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (reps * iters) * 1e3


out = {"batch": bs}
for on in ((True,) if os.environ.get("AB_ONLY") == "stack" else (False, True)):
    stack.set_enabled(on)
    torch.manual_seed(0)
    tr = Trainer(dict(hp), device=dev, sync_losses=False).to(dev)
    model = tr.model
    x = ops.transpose_ct(data[0])
    n = hp["num_layers"]
    with torch.no_grad():
        _, zs = model.enc(x, needed={0, n - 1})
        z_list = [None] * n
        z_list[0] = zs[0][:, :, :model.shallow_latent_d].contiguous()
        z_list[n - 1] = zs[n - 1][:, :, :model.latent_d].contiguous()
        enc_us = graph_time(lambda: model.enc(x, needed=set()))
        dec_us = graph_time(lambda: model.dec(z_list))
    tr.enable_cuda_graph(data, hp, 0, warmup=3)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(10):
        tr.gen_update(data, hp, 0)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(200):
        tr.gen_update(data, hp, 0)
    e1.record()
    torch.cuda.synchronize()
    out["stack" if on else "per_layer"] = {"enc_fwd_us": round(enc_us, 1), "dec_fwd_us": round(dec_us, 1),
                                           "step_us": round(e0.elapsed_time(e1) / 200 * 1e3, 1), "launches_per_step": tr.launches_per_step}
    tr.gen_opt.close()
    ops.unregister_grad_buffers()
    del tr
print(json.dumps(out))
