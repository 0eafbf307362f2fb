"""Step times of the other BASELINE configs on one GPU (parity for them is in tests/test_gpu_parity.py):
   config 2: configs/len8_data_aug_hm_vae.yaml training step, B=8 (launch-overhead bound: K=3, T=8)
   config 5: configs/trajectory_model.yaml training step, B=8, T=128, K=31
Both through Trainer.gen_update, eager and as a CUDA graph.  Prints one JSON object."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import yaml  # noqa: E402

from hm_vae_b200 import ops  # noqa: E402
from hm_vae_b200.trainer_motion_vae import Trainer  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)


def timed(fn, steps):
    for _ in range(5):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def run(cfg_name, bs, make_data, steps=200):
    hp = yaml.safe_load(open(os.path.join(ROOT, "configs", cfg_name)))
    torch.manual_seed(0)
    tr = Trainer(dict(hp), device=dev, sync_losses=False).to(dev)
    data = make_data(hp, bs)
    eager = timed(lambda: tr.gen_update(data, hp, 0), 50)
    tr.enable_cuda_graph(data, hp, 0, warmup=3)
    graph = timed(lambda: tr.gen_update(data, hp, 0), steps)
    out = tr.gen_update(data, hp, 0)
    res = {"config": cfg_name, "batch": bs, "eager_ms_per_step": eager, "graph_ms_per_step": graph,
           "sequences_per_s": bs / (graph * 1e-3), "launches_per_step": tr.launches_per_step, "loss": float(out[0]),
           "trainable_params": sum(p.numel() for p in tr.model.parameters() if p.requires_grad)}
    ops.unregister_grad_buffers()
    return res


def hmvae_data(hp, bs):
    T = hp["train_seq_len"]
    g = torch.Generator().manual_seed(1234)
    rot = ops.rot6d_to_rotmat(torch.randn(bs, T, 24, 6, generator=g).to(dev))
    return (torch.stack((rot[..., 0], rot[..., 1]), dim=-2).reshape(bs, T, -1).contiguous(), rot.reshape(bs, T, -1).contiguous())


def traj_data(hp, bs):
    T = hp["train_seq_len"]
    g = torch.Generator().manual_seed(1234)
    d6, dm = hmvae_data(hp, bs)
    return (d6, dm, None, torch.randn(bs, T, 72, generator=g).to(dev), None, None, torch.randn(bs, T, 3, generator=g).to(dev))


print(json.dumps({"len8": run("len8_data_aug_hm_vae.yaml", 8, hmvae_data), "trajectory": run("trajectory_model.yaml", 8, traj_data)}))
