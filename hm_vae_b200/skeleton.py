"""Drop-in SkeletonConv / SkeletonPool / SkeletonUnpool and the skeleton topology helpers.

Same constructor signatures, attribute names, parameter names/shapes/requires_grad flags and state_dict order as the
reference (skeleton.py:9-105, 159-261, 306-315, 364-411), so reference checkpoints load unchanged.  ``forward`` runs
the sm_100a kernels of libhmvae_b200.so: the weight mask is never multiplied in -- masked (j_out, j_in) blocks are
simply not visited -- and the pool/unpool matrices are never multiplied either (they are index gathers).  The dense
``mask`` / ``weight`` parameters of pool/unpool exist only for state_dict compatibility.
"""
import json
import math

import numpy as np
import torch
import torch.nn as nn

from . import ops


# ------------------------------------------------------------------------------------------------ topology (ints)
def get_edges(parent_json):
    """skeleton.py:306-315.  Accepts a json path or a parents list; the virtual root edge (0, J) comes first."""
    parents = json.load(open(parent_json, "r")) if isinstance(parent_json, str) else list(parent_json)
    return [(0, len(parents))] + [(int(parents[i]), i) for i in range(1, len(parents))]


def calc_edge_mat(edges):
    """skeleton.py:364-387: all-pairs edge distance (min-plus closure of the share-a-joint adjacency).

    The reference's adjacency pass also matches every edge with itself, so the diagonal is 1 (not 0); kept."""
    n = len(edges)
    ends = np.asarray([[e[0], e[1]] for e in edges], dtype=np.int64)
    share = (ends[:, None, :, None] == ends[None, :, None, :]).any(axis=(2, 3))
    dist = np.where(share, 1, 100000).astype(np.int64)
    for k in range(n):
        dist = np.minimum(dist, dist[:, k:k + 1] + dist[k:k + 1, :])
    return dist.tolist()


def find_neighbor(edges, d):
    """skeleton.py:390-411."""
    mat = np.asarray(calc_edge_mat(edges))
    return [np.nonzero(row <= d)[0].tolist() for row in mat]


def _chains(edges):
    """Edge-index chains from the root to branch joints / end effectors (find_seq, skeleton.py:180-195)."""
    degree = {}
    for a, b in edges:
        degree[a] = degree.get(a, 0) + 1
        degree[b] = degree.get(b, 0) + 1
    children = {}
    for idx, (a, b) in enumerate(edges):
        children.setdefault(a, []).append((idx, b))
    out = []

    def walk(joint, chain):
        if degree.get(joint, 0) > 2 and joint != 0:
            out.append(chain)
            chain = []
        if degree.get(joint, 0) == 1:
            out.append(chain)
            return
        for idx, child in children.get(joint, []):
            walk(child, chain + [idx])

    walk(0, [])
    return out


# ------------------------------------------------------------------------------------------------ modules
class SkeletonConv(nn.Module):
    def __init__(self, neighbour_list, in_channels, out_channels, kernel_size, joint_num, stride=1, padding=0,
                 bias=True, padding_mode='zeros', add_offset=False, in_offset_channel=0):
        self.in_channels_per_joint = in_channels // joint_num
        self.out_channels_per_joint = out_channels // joint_num
        if in_channels % joint_num != 0 or out_channels % joint_num != 0:
            raise Exception('BAD')
        super(SkeletonConv, self).__init__()
        if add_offset:
            raise NotImplementedError("add_offset=True (SkeletonLinear offset encoder) is never used by hm-vae and is not built")
        if padding_mode == 'zeros':
            padding_mode = 'constant'
        if padding_mode == 'reflection':
            padding_mode = 'reflect'
        if padding_mode not in ('constant', 'reflect'):
            raise Exception('Unsupported padding mode {}'.format(padding_mode))
        self.neighbour_list = neighbour_list
        self.add_offset = add_offset
        self.joint_num = joint_num
        self.kernel_size = kernel_size
        self.stride, self.dilation, self.groups = stride, 1, 1
        self.padding, self.padding_mode = padding, padding_mode
        self._padding_repeated_twice = (padding, padding)
        ci = self.in_channels_per_joint
        self.expanded_neighbour_list = [[k * ci + i for k in nb for i in range(ci)] for nb in neighbour_list]
        self.expanded_neighbour_list_offset = []

        if not bias:
            self.register_parameter('bias', None)
        mask = torch.zeros(out_channels, in_channels, kernel_size)
        co = self.out_channels_per_joint
        for j, cols in enumerate(self.expanded_neighbour_list):
            mask[co * j: co * (j + 1), cols, :] = 1
        self.mask = nn.Parameter(mask, requires_grad=False)          # registered before weight, as in the reference
        self.weight = nn.Parameter(torch.zeros(out_channels, in_channels, kernel_size))
        if bias:
            self.bias = nn.Parameter(torch.zeros(out_channels))
        self.description = 'SkeletonConv(in_channels_per_armature={}, out_channels_per_armature={}, kernel_size={}, ' \
                           'joint_num={}, stride={}, padding={}, bias={})'.format(
                               in_channels // joint_num, out_channels // joint_num, kernel_size, joint_num, stride, padding, bias)
        self._plans = {}
        self.reset_parameters()

    def reset_parameters(self):
        """Per output joint: kaiming_uniform(a=sqrt 5) inside the neighbour block, zeros elsewhere; bias U(+-1/sqrt(fan_in))
        per joint block (skeleton.py:70-89) -- same RNG consumption order as the reference."""
        co = self.out_channels_per_joint
        with torch.no_grad():
            self.weight.zero_()
            for j, cols in enumerate(self.expanded_neighbour_list):
                block = torch.empty(co, len(cols), self.kernel_size)
                nn.init.kaiming_uniform_(block, a=math.sqrt(5))
                self.weight[co * j: co * (j + 1), cols, :] = block.to(self.weight.device)
                if self.bias is not None:
                    bound = 1 / math.sqrt(len(cols) * self.kernel_size)
                    b = torch.empty(co)
                    nn.init.uniform_(b, -bound, bound)
                    self.bias[co * j: co * (j + 1)] = b.to(self.bias.device)

    def set_offset(self, offset):
        raise Exception('Wrong Combination of Parameters')

    def plan(self, **fused):
        key = (torch.cuda.current_device(),
               tuple(sorted((k, tuple(v) if isinstance(v, (list, tuple)) else v) for k, v in fused.items())))
        p = self._plans.get(key)
        if p is None:
            p = ops.ConvPlan(self.neighbour_list, self.in_channels_per_joint, self.out_channels_per_joint, self.kernel_size,
                             self.stride, self.padding, self.padding_mode, **fused)
            self._plans[key] = p
        p.exact = bool(getattr(self, "exact", False))     # per-layer precision override: fp32 CUDA-core kernels instead of TF32
        return p

    def forward(self, input):
        return ops.skeleton_conv(input, self.weight, self.bias, self.plan())

    def fused_forward(self, input, **fused):
        """forward with a fused prologue (upsample / unpool_src+src_joints) and/or epilogue (lrelu, output layout)."""
        return ops.skeleton_conv(input, self.weight, self.bias, self.plan(**fused))

    def __deepcopy__(self, memo):
        import copy
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            setattr(new, k, {} if k == '_plans' else copy.deepcopy(v, memo))
        return new


class SkeletonPool(nn.Module):
    def __init__(self, edges, pooling_mode, channels_per_edge, last_pool=False):
        super(SkeletonPool, self).__init__()
        if pooling_mode != 'mean':
            raise Exception('Unimplemented pooling mode in matrix_implementation')
        self.channels_per_edge = channels_per_edge
        self.pooling_mode = pooling_mode
        self.edge_num = len(edges)
        self.seq_list = _chains(edges)
        self.pooling_list = []
        self.new_edges = []
        for seq in self.seq_list:
            if last_pool:
                self.pooling_list.append(seq)
                continue
            rest = seq
            if len(rest) % 2 == 1:
                self.pooling_list.append([rest[0]])
                self.new_edges.append(edges[rest[0]])
                rest = rest[1:]
            for i in range(0, len(rest), 2):
                self.pooling_list.append([rest[i], rest[i + 1]])
                self.new_edges.append([edges[rest[i]][0], edges[rest[i + 1]][1]])
        self.description = 'SkeletonPool(in_edge_num={}, out_edge_num={})'.format(len(edges), len(self.pooling_list))
        c = channels_per_edge
        weight = torch.zeros(len(self.pooling_list) * c, self.edge_num * c)
        eye = torch.arange(c)
        for i, pair in enumerate(self.pooling_list):
            for j in pair:
                weight[i * c + eye, j * c + eye] = 1.0 / len(pair)
        self.weight = nn.Parameter(weight, requires_grad=False)       # state_dict compatibility only

    def forward(self, input: torch.Tensor, lrelu=False):
        return ops.skeleton_pool(input, self.pooling_list, self.channels_per_edge, lrelu)


class SkeletonUnpool(nn.Module):
    def __init__(self, pooling_list, channels_per_edge):
        super(SkeletonUnpool, self).__init__()
        self.pooling_list = pooling_list
        self.input_edge_num = len(pooling_list)
        self.output_edge_num = sum(len(t) for t in pooling_list)
        self.channels_per_edge = channels_per_edge
        self.description = 'SkeletonUnpool(in_edge_num={}, out_edge_num={})'.format(self.input_edge_num, self.output_edge_num)
        c = channels_per_edge
        weight = torch.zeros(self.output_edge_num * c, self.input_edge_num * c)
        eye = torch.arange(c)
        self.src = [0] * self.output_edge_num
        for i, pair in enumerate(self.pooling_list):
            for j in pair:
                weight[j * c + eye, i * c + eye] = 1
                self.src[j] = i
        self.weight = nn.Parameter(weight)                            # state_dict compatibility only
        self.weight.requires_grad_(False)

    def forward(self, input: torch.Tensor):
        return ops.skeleton_unpool(input, self.src, self.channels_per_edge, self.input_edge_num)
