"""ORACLE (test infrastructure only -- never imported by the product path): CPU restatement of the reference's batch assembly,
the step right before the hot path.

Follows /root/reference/utils_motion_vae.py:
  rand_rotation_matrix   :17-57   (Graphics Gems III random rotation from three uniform numbers, float64)
  change_fps             :65-81   (random temporal sub-sampling factor; returns the factor instead of the sliced array)
  MotionSeqData.__getitem__ :124-187 (579-column layout, standardisation, random root rotation, 6D re-derivation)
Pinned by tests/golden/batch.npz, which oracle/make_golden_batch.py produced by running the real reference.
"""
import numpy as np

N_DIM = 579
SL_6D, SL_MAT, SL_POS = slice(0, 144), slice(144, 360), slice(360, 432)
SL_LINV, SL_ANGV, SL_ROOTV = slice(432, 504), slice(504, 576), slice(576, 579)


def rand_rotation_matrix(deflection=1.0, randnums=None):
    """utils_motion_vae.py:17-57.  randnums: three numbers in [0, 1]."""
    theta, phi, z = np.asarray(randnums, dtype=np.float64)
    theta = theta * 2.0 * deflection * np.pi
    phi = phi * 2.0 * np.pi
    z = z * 2.0 * deflection
    r = np.sqrt(z)
    V = np.array([np.sin(phi) * r, np.cos(phi) * r, np.sqrt(2.0 - z)])
    st, ct = np.sin(theta), np.cos(theta)
    R = np.array(((ct, st, 0.0), (-st, ct, 0.0), (0.0, 0.0, 1.0)))
    return (np.outer(V, V) - np.eye(3)).dot(R)


def change_fps_factor(n_frames, train_seq_len, draw):
    """utils_motion_vae.py:65-81.  `draw()` returns one element of [1, 2, 3, 4, 5, 6, 8, 10, 12]; up to 10 tries, the first
    factor that leaves at least train_seq_len frames wins, else 1 (the original data)."""
    for _ in range(10):
        f = draw()
        if len(range(0, n_frames, f)) >= train_seq_len:
            return f
    return 1


def assemble(window, mean_std, root_rot=None):
    """utils_motion_vae.py:140-187 for ONE cropped window [T, 579] (float32).  mean_std: [2, 579] with the zero stds already
    replaced by 1 (:104).  root_rot: 3x3 float64 from rand_rotation_matrix, or None (random_root_rot_flag off).
    Returns the 7-tuple of float32 arrays."""
    window = np.asarray(window, dtype=np.float32)
    std = ((window - mean_std[0][None, :]) / mean_std[1][None, :]).astype(np.float32)      # numpy promotes to the file's dtype
    rot6d = window[:, SL_6D].copy()
    rotmat = window[:, SL_MAT].copy()
    rot_pos = window[:, SL_POS].copy()
    joint_pos, linear_v, angular_v, root_v = std[:, SL_POS], std[:, SL_LINV], std[:, SL_ANGV], std[:, SL_ROOTV]
    if root_rot is not None:
        M = np.asarray(root_rot, dtype=np.float64).astype(np.float32)                      # torch.from_numpy(...).float()
        T = window.shape[0]
        root = rotmat[:, :9].reshape(T, 3, 3)
        rotmat[:, :9] = np.einsum("ij,tjk->tik", M, root).reshape(T, 9)                    # fp32 matmul
        aug_v = np.einsum("ij,tj->ti", M, window[:, SL_ROOTV])                             # fp32
        root_v = ((aug_v - mean_std[0][None, SL_ROOTV]) / mean_std[1][None, SL_ROOTV]).astype(np.float32)
        R = rotmat.reshape(T, 24, 3, 3)
        rot6d = np.stack((R[..., 0], R[..., 1]), axis=-2).reshape(T, 144)                  # columns 0 and 1 of every joint
    return rot6d, rotmat, rot_pos, joint_pos, linear_v, angular_v, root_v
