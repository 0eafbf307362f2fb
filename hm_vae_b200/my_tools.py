"""Drop-in rotation transforms: my_tools.rotation_matrix_from_ortho6d (my_tools.py:19-39) and
torchgeometry.angle_axis_to_rotation_matrix (call sites seq_two_hier_sa_vae.py:650, trajectory_pred_model.py:451)."""
import torch

from . import ops


def rotation_matrix_from_ortho6d(poses):
    """poses [..., 6] -> [..., 3, 3]; columns are (x, y, z) with x = normalize(a), z = normalize(x X b), y = z X x."""
    return ops.rot6d_to_rotmat(poses)


def rotation_matrix_to_ortho6d(rot):
    """[..., 3, 3] -> [..., 6] (first two columns); inlined everywhere in the reference, e.g. seq_two_hier_sa_vae.py:666-667."""
    return torch.stack((rot[..., 0], rot[..., 1]), dim=-2).reshape(*rot.shape[:-2], 6)


def angle_axis_to_rotation_matrix(angle_axis):
    """[N, 3] -> [N, 4, 4] homogeneous rotation (callers slice [:, :3, :3])."""
    return ops.angle_axis_to_rotation_matrix(angle_axis)
