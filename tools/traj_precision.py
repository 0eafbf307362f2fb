"""Which conv layer of the trajectory model (K=31, T=128) breaks the 2e-3 loss tolerance under TF32?  Runs the golden step with
every subset of layers forced to the fp32 CUDA-core kernels (SkeletonConv.exact) and prints the loss errors + step time."""
import itertools
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from hm_vae_b200 import ops  # noqa: E402
from hm_vae_b200.trajectory_pred_model import TrajectoryModel  # noqa: E402
from oracle import hmvae_ref as O  # noqa: E402
from test_oracle_golden import HPT  # noqa: E402

DEV = "cuda"
g = dict(np.load(os.path.join(ROOT, "tests", "golden", "models.npz")))
smpl = dict(np.load(os.path.join(ROOT, "hm_vae_b200", "data", "smpl24.npz")))
parents, off = smpl["parents"].tolist(), torch.from_numpy(smpl["offsets"])
ms = torch.from_numpy(smpl["mean_std"])
ora = O.TrajectoryOracle(HPT, ms, parents).init(seed=0)
model = TrajectoryModel(dict(HPT), device=DEV)
sd = model.state_dict()
for k, v in ora.params.items():
    sd[k] = v.detach().clone()
model.load_state_dict(sd)
model = model.to(DEV)
batch = O.synthetic_batch(2, 128, parents, off, seed=1234, mean_std=ms)
data = (batch["seq_rot_6d"], batch["seq_rot_mat"], batch["seq_rot_pos"], batch["seq_joint_pos"], None, None, batch["seq_root_v"])
ref = g["traj_losses"]
for r in range(5):
    for subset in itertools.combinations(range(4), r):
        for i, c in enumerate(model.enc.convs):
            c.exact = i in subset
        for p in model.parameters():
            p.grad = None
        res = model(data, HPT, 0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            model(data, HPT, 0)
        e1.record()
        torch.cuda.synchronize()
        got = np.asarray([float(res[0]), float(res[6]), float(res[8])])
        gerr = max(abs(float((p.grad.double() ** 2).sum()) - g["traj_grad/" + k][2]) / g["traj_grad/" + k][2]
                   for k, p in model.named_parameters() if p.requires_grad)
        print("exact layers %-12s loss rel err (total, root_v, root_trans) %s  worst grad sq-sum rel err %.2e  eager ms/step %.3f" % (
            subset, np.array2string(np.abs(got - ref) / np.abs(ref), precision=2), gerr, e0.elapsed_time(e1) / 5), flush=True)
