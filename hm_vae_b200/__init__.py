"""hm_vae_b200 -- B200-native (sm_100a) implementation of the hm-vae hot path behind the reference's module surface.

Importing the package loads libhmvae_b200.so (hand-written CUDA behind a C ABI, include/hmvae_b200.h); it raises if the
library is missing -- there is no CPU or eager-PyTorch fallback.
"""
from . import _lib  # noqa: F401  (loads the shared library; fails loudly)
from .fk_layer import ForwardKinematicsLayer
from .my_tools import angle_axis_to_rotation_matrix, rotation_matrix_from_ortho6d
from .skeleton import SkeletonConv, SkeletonPool, SkeletonUnpool, calc_edge_mat, find_neighbor, get_edges

__all__ = ["ForwardKinematicsLayer", "SkeletonConv", "SkeletonPool", "SkeletonUnpool", "angle_axis_to_rotation_matrix",
           "calc_edge_mat", "find_neighbor", "get_edges", "rotation_matrix_from_ortho6d"]
