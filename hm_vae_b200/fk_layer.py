"""Drop-in ForwardKinematicsLayer (fk_layer.py:11-93) over the sm_100a FK kernels.

``forward(rotations[N,J,3,3] | [N,J,6], positions=None) -> [N,J,3]``; fwd and bwd are one kernel each instead of
23 gather+bmm+slice-copy triplets.  Unlike the reference (whose default constructor reads absolute paths on the
author's machine, fk_layer.py:18-19), the default skeleton is the SMPL-24 fixture shipped in ``data/smpl24.npz``.
"""
import os

import numpy as np
import torch

from . import ops

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "smpl24.npz")


def load_smpl24():
    d = np.load(_DATA)
    return d["parents"].astype(np.int64).tolist(), d["offsets"].astype(np.float32), d["mean_std"]


class ForwardKinematicsLayer(torch.nn.Module):
    """ Forward Kinematics Layer Class """

    def __init__(self, device=torch.device("cuda"), parents=None, positions=None):
        super().__init__()
        if parents is None and positions is None:
            parents, positions, _ = load_smpl24()
        self.device = device
        self._parents_list = [int(p) for p in parents]
        if any(p >= i for i, p in enumerate(self._parents_list) if i > 0):
            raise ValueError("parents[i] must be < i (the reference's joint loop assumes it, fk_layer.py:76-78)")
        # plain tensor attributes, as in the reference (fk_layer.py:25-26): not buffers, not in the state_dict
        self.parents = torch.tensor(self._parents_list, dtype=torch.long, device=device)
        self.positions = torch.from_numpy(np.asarray(positions)).float()[None, :, :].to(device)   # 1 X J X 3

    def forward(self, rotations, positions=None):
        if not ((rotations.shape[-1] == 3 and rotations.shape[-2] == 3) or rotations.shape[-1] == 6):
            raise ValueError("rotations must be [N,J,3,3] or [N,J,6]")
        if rotations.shape[-1] == 6:
            rot = rotations.reshape(rotations.shape[0], rotations.shape[1], 6)
        else:
            rot = rotations.reshape(rotations.shape[0], rotations.shape[1], 3, 3)
        offsets = self.positions[0].contiguous()
        if offsets.device != rot.device:
            offsets = offsets.to(rot.device)
        return ops.forward_kinematics(rot, offsets, positions, self._parents_list)
