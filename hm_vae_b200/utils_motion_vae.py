"""Host mirror of the reference's batch assembly (utils_motion_vae.py) over the B200 kernels.

Same names and meaning as the reference: ``rand_rotation_matrix(deflection, randnums)`` (:17-57), ``change_fps(ori_data,
train_seq_len)`` (:65-81) and the 7-tuple layout of ``MotionSeqData.__getitem__`` (:124-187).  What moves to the device is
the arithmetic: ``DeviceBatchAssembler`` takes a batch of already cropped ``[B, T, 579]`` windows (pinned host or device) and
produces the seven training tensors with ONE kernel (``hmvae_batch_assemble``): slicing, float64 standardisation, random
root rotation (one rotation per sequence, ``hmvae_rand_rotation``), 6D re-derivation.  The reference does this per sequence in
one DataLoader worker and ships seven tensors host -> device every step (five of them unused by HM-VAE training).

The random draws (crop offset, fps factor, the three uniform numbers per sequence) stay on the host and use the same global
``random`` / ``np.random`` generators in the same order as the reference, so a seeded run selects the same windows.
"""
import random

import numpy as np
import torch

from . import _lib
from ._lib import check, lib, stream

N_DIM = 579
FPS_FACTORS = [1, 2, 3, 4, 5, 6, 8, 10, 12]


def rand_rotation_matrix(deflection=1.0, randnums=None):
    """utils_motion_vae.py:17-57 -- one 3x3 float64 matrix on the host (numpy), for callers that used the reference function."""
    if randnums is None:
        randnums = np.random.uniform(size=(3,))
    theta, phi, z = randnums
    theta = theta * 2.0 * deflection * np.pi
    phi = phi * 2.0 * np.pi
    z = z * 2.0 * deflection
    r = np.sqrt(z)
    V = np.array((np.sin(phi) * r, np.cos(phi) * r, np.sqrt(2.0 - z)))
    st, ct = np.sin(theta), np.cos(theta)
    R = np.array(((ct, st, 0), (-st, ct, 0), (0, 0, 1)))
    return (np.outer(V, V) - np.eye(3)).dot(R)


def rand_rotation_matrices(randnums, deflection=1.0, device="cuda"):
    """Batched device version: randnums [B, 3] (float64, in [0, 1]) -> [B, 3, 3] float32, computed in float64 on the GPU."""
    rn = torch.as_tensor(np.asarray(randnums, dtype=np.float64)).reshape(-1, 3).to(device).contiguous()
    out = torch.empty((rn.shape[0], 3, 3), device=rn.device, dtype=torch.float32)
    check(lib.hmvae_rand_rotation(rn.data_ptr(), float(deflection), out.data_ptr(), rn.shape[0], stream()), "rand_rotation")
    return out


def change_fps(ori_data, train_seq_len):
    """utils_motion_vae.py:65-81: up to 10 random sub-sampling factors; the first one that leaves >= train_seq_len frames."""
    res, tries = ori_data, 0
    while tries < 10:
        f = random.sample(FPS_FACTORS, 1)[0]
        dest = ori_data[0::f, :]
        tries += 1
        if dest.shape[0] >= train_seq_len:
            res = dest
            break
    return res


def crop_window(ori_pose_seq_data, train_seq_len, fps_aug_flag=False):
    """The host-side index logic of ``__getitem__`` (:128-141): optional fps change, then a random window of train_seq_len
    frames.  Returns the [T, 579] window, or None when the sequence is too short (the reference then draws another index)."""
    if fps_aug_flag:
        ori_pose_seq_data = change_fps(ori_pose_seq_data, train_seq_len)
    timesteps = ori_pose_seq_data.shape[0]
    if train_seq_len > timesteps:
        return None
    t0 = random.sample(list(range(timesteps - train_seq_len + 1)), 1)[0]
    return ori_pose_seq_data[t0:t0 + train_seq_len, :]


class DeviceBatchAssembler:
    """``assembler(raw_windows) -> (seq_rot_6d, seq_rot_mat, seq_rot_pos, seq_joint_pos, seq_linear_v, seq_angular_v,
    seq_root_v)``, the data tuple ``TwoHierSAVAEModel.forward`` / ``TrajectoryModel.forward`` take."""

    def __init__(self, mean_std, random_root_rot_flag=False, device="cuda", only_hmvae_inputs=False):
        ms = np.array(mean_std, dtype=np.float64, copy=True)
        if ms.shape != (2, N_DIM):
            raise ValueError("mean_std must be [2, %d]" % N_DIM)
        ms[1, ms[1, :] == 0] = 1.0                                   # :104
        self.device = torch.device(device)
        self.mean = torch.from_numpy(ms[0]).to(self.device).contiguous()
        self.std = torch.from_numpy(ms[1]).to(self.device).contiguous()
        self.random_root_rot_flag = bool(random_root_rot_flag)
        self.only_hmvae_inputs = bool(only_hmvae_inputs)              # HM-VAE training reads only the first two tensors

    def __call__(self, raw, randnums=None):
        """raw: [B, T, 579] float32 (device, or host -- copied).  randnums: [B, 3] uniform numbers for the root rotations
        (drawn from np.random in batch order when the augmentation is on and none are given)."""
        raw = torch.as_tensor(raw)
        if raw.dim() != 3 or raw.shape[-1] != N_DIM:
            raise ValueError("raw must be [B, T, %d], got %s" % (N_DIM, tuple(raw.shape)))
        raw = raw.to(device=self.device, dtype=torch.float32, non_blocking=True).contiguous()
        b, t, _ = raw.shape
        rot = None
        if self.random_root_rot_flag:
            if randnums is None:
                randnums = np.stack([np.random.uniform(size=(3,)) for _ in range(b)])
            rot = rand_rotation_matrices(randnums, device=self.device).reshape(b, 9)
        mk = lambda c: torch.empty((b, t, c), device=self.device, dtype=torch.float32)
        outs = [mk(144), mk(216)] + ([None] * 5 if self.only_hmvae_inputs else [mk(72), mk(72), mk(72), mk(72), mk(3)])
        ptr = lambda x: x.data_ptr() if x is not None else None
        check(lib.hmvae_batch_assemble(raw.data_ptr(), ptr(rot), self.mean.data_ptr(), self.std.data_ptr(), b, t,
                                       *[ptr(o) for o in outs], stream()), "batch_assemble")
        return tuple(outs)
