"""Root-trajectory regressor over the B200 kernels (trajectory_pred_model.py:45-115, 174-260, 289-303).

Encoder = 4 x (SkeletonConv K=31 stride 1 -> SkeletonPool -> LeakyReLU); ``fc_mapping`` per frame; the Python loop
of ``gen_motion_w_trajectory`` (127 tiny launches, twice per step) and both MSEs are one prefix-scan kernel
(``hmvae_traj_fwdbwd``) that also emits the gradient w.r.t. the predicted root velocity.
"""
import torch
import torch.nn as nn

from . import ops
from .fk_layer import ForwardKinematicsLayer, load_smpl24
from .skeleton import SkeletonConv, SkeletonPool, find_neighbor, get_edges


class Encoder(nn.Module):
    def __init__(self, args, topology):
        super(Encoder, self).__init__()
        self.topologies = [topology]
        self.channel_base = [3] if args['trajectory_input_joint_pos'] else [6]
        self.channel_list = []
        self.edge_num = [len(topology)]
        self.pooling_list = []
        self.layers = nn.ModuleList()
        self.args = args
        self.convs, self.pools = [], []
        n = args['num_layers']
        kernel_size = args['kernel_size']
        padding = (kernel_size - 1) // 2
        for i in range(n):
            self.channel_base.append(self.channel_base[-1] * 2)
        for i in range(n):
            neighbor_list = find_neighbor(self.topologies[i], args['skeleton_dist'])
            in_channels = self.channel_base[i] * self.edge_num[i]
            out_channels = self.channel_base[i + 1] * self.edge_num[i]
            if i == 0:
                self.channel_list.append(in_channels)
            self.channel_list.append(out_channels)
            seq = []
            for _ in range(args['extra_conv']):
                seq.append(SkeletonConv(neighbor_list, in_channels=in_channels, out_channels=in_channels,
                                        joint_num=self.edge_num[i], kernel_size=kernel_size, stride=1, padding=padding,
                                        padding_mode=args['padding_mode'], bias=True))
            seq.append(SkeletonConv(neighbor_list, in_channels=in_channels, out_channels=out_channels,
                                    joint_num=self.edge_num[i], kernel_size=kernel_size, stride=1, padding=padding,
                                    padding_mode=args['padding_mode'], bias=True))
            self.convs.append(seq[-1])
            pool = SkeletonPool(edges=self.topologies[i], pooling_mode=args['skeleton_pool'],
                                channels_per_edge=out_channels // len(neighbor_list), last_pool=(i == n - 1))
            self.pools.append(pool)
            seq.append(pool)
            seq.append(nn.LeakyReLU(negative_slope=0.2))
            self.layers.append(nn.Sequential(*seq))
            self.topologies.append(pool.new_edges)
            self.pooling_list.append(pool.pooling_list)
            self.edge_num.append(len(self.topologies[-1]))

    def forward(self, input, offset=None):
        for i in range(len(self.layers)):
            for m in list(self.layers[i])[:-3]:
                input = m(input)
            conv, pool = self.convs[i], self.pools[i]
            if all(len(p) == 1 and p[0] == k for k, p in enumerate(pool.pooling_list)):
                input = conv.fused_forward(input, lrelu=True)
            else:
                input = pool(conv(input), lrelu=True)
        return input


class TrajectoryModel(nn.Module):
    def __init__(self, hp, parent_json=None, device=None):
        super(TrajectoryModel, self).__init__()
        self.latent_d = hp['latent_d']
        self.n_joints = hp['n_joints']
        self.input_dim = hp['input_dim']
        self.output_dim = hp['output_dim']
        self.max_timesteps = hp['train_seq_len']
        parents, offsets, mean_std = load_smpl24()
        edges = get_edges(parent_json if parent_json is not None else parents)
        dev = torch.device("cuda") if device is None else torch.device(device)
        self.fk_layer = ForwardKinematicsLayer(device=dev)
        self.hp = hp
        self.enc = Encoder(hp, edges)
        self.d_model = self.enc.channel_base[-1]
        self.fc_mapping = nn.Linear(self.d_model * 7, 3)
        mean_std = mean_std.copy()
        mean_std[1, mean_std[1, :] == 0] = 1.0
        self._mean3 = [float(v) for v in mean_std[0, 576:579]]
        self._std3 = [float(v) for v in mean_std[1, 576:579]]
        self.mean_vals = torch.from_numpy(mean_std[0, :]).float()[None, :].to(dev)
        self.std_vals = torch.from_numpy(mean_std[1, :]).float()[None, :].to(dev)

    def _to_device(self, t):
        return t.to(device=self.mean_vals.device, dtype=torch.float32, non_blocking=True)

    def forward(self, data, hp, iterations, multigpus=False, validation_flag=False):
        seq_rot_6d, seq_joint_pos, seq_root_v = data[0], data[3], data[6]
        seq_root_v = self._to_device(seq_root_v).contiguous()          # bs X T X 3, standardised
        if hp['trajectory_input_joint_pos']:
            encoder_input = self._to_device(seq_joint_pos).contiguous()  # bs X T X (24*3), standardised
        else:
            encoder_input = self._to_device(seq_rot_6d).contiguous()
        bs, timesteps, _ = encoder_input.size()
        dev = encoder_input.device
        latent = self.enc(ops.transpose_ct(encoder_input))             # bs X (7*d) X T
        feat = ops.transpose_ct(latent)                                # bs X T X (7*d)   (edge-major, channel-minor)
        root_v_out = ops.linear(feat, self.fc_mapping.weight, self.fc_mapping.bias)   # bs X T X 3

        sums = torch.zeros(2, device=dev, dtype=torch.float32)
        w_t = hp['rec_root_trans_w'] if hp['use_accumulation_root_v'] else 0.0
        d_root_v = ops.traj_fwdbwd(root_v_out.detach(), seq_root_v, self._mean3, self._std3, self.n_joints,
                                   hp['rec_root_v_w'], w_t, sums, want_grad=not validation_flag)
        l_rec_root_v = sums[0] / (bs * timesteps * 3)
        if hp['use_accumulation_root_v']:
            l_rec_root_trans = sums[1] / (timesteps * bs * self.n_joints * 3)
        else:
            l_rec_root_trans = torch.zeros(1, device=dev)
        l_total = hp['rec_root_v_w'] * l_rec_root_v + hp['rec_root_trans_w'] * l_rec_root_trans
        if not validation_flag:
            with ops.wgrad_overlap():
                root_v_out.backward(d_root_v)
        zero = torch.zeros(1, device=dev)
        return l_total, zero, zero, zero, zero, zero, l_rec_root_v, zero, l_rec_root_trans

    def l2_criterion(self, pred, gt):
        return ops.l2_criterion(pred, gt)

    def de_standardize(self, output_data, start_idx, end_idx):
        if output_data.dim() == 2:
            return self.mean_vals[:, start_idx:end_idx] + self.std_vals[:, start_idx:end_idx] * output_data
        return self.mean_vals[None][:, :, start_idx:end_idx] + self.std_vals[None][:, :, start_idx:end_idx] * output_data

    def gen_motion_w_trajectory(self, pose_data, root_v_data, need_destandardize=True):
        """pose_data T X bs X 24 X 3 (root at origin), root_v_data T X bs X 3 -> absolute poses (prefix sum over T)."""
        v = self.de_standardize(root_v_data, 576, 579) if need_destandardize else root_v_data
        v = torch.cat([torch.zeros_like(v[:1]), v[1:]], dim=0)
        return pose_data + torch.cumsum(v, dim=0)[:, :, None, :]
