"""Bandwidth of hmvae_batch_assemble (algorithmic bytes: 579*4 in + 651*4 out per frame) against the measured copy peak."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from hm_vae_b200 import utils_motion_vae as U  # noqa: E402
from hm_vae_b200.fk_layer import load_smpl24  # noqa: E402

dev = torch.device("cuda", 0)
_, _, mean_std = load_smpl24()
res = {}
for b in (32, 4096):
    raw = torch.randn(b, 64, 579, device=dev)
    rn = np.random.RandomState(0).uniform(size=(b, 3))
    asm = U.DeviceBatchAssembler(mean_std, random_root_rot_flag=True, device=dev)
    for _ in range(3):
        asm(raw, randnums=rn)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        asm(raw, randnums=rn)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    nbytes = b * 64 * (579 + 651) * 4
    res[str(b)] = {"ms_per_call_incl_rotation_and_allocs": ms, "gbs": nbytes / (ms * 1e-3) / 1e9, "frac": nbytes / (ms * 1e-3) / 1e9 / bench.peaks()["hbm"]}
print(json.dumps(res))
