"""HM-VAE networks over the B200 kernels: Encoder / Decoder / TwoHierSAVAEModel.

Mirrors the reference's module surface (seq_two_hier_sa_vae.py:53-474, 560-639): constructor arguments, attribute
names, the 74-key state_dict (including the duplicated ``dec.enc.*`` keys and the frozen pool/unpool/mask tensors),
the 10-tuple returned by ``forward`` and the fact that ``forward`` itself runs the backward pass.

What differs is how a step executes:
  * SkeletonConv never multiplies by the mask; reflect padding, the decoder's x2 upsample and the unpool gather are
    index arithmetic in the conv loader; LeakyReLU (and pooling) are conv / pool epilogues.
  * ``z_vec_list[1:3]`` / ``hier_feats[1:3]`` never influence the output (reference :278-288), so they are not computed.
  * GT FK, rot6d->R, FK, the three MSEs and their whole backward are ONE kernel (``hmvae_recon_fwdbwd``); the backward
    pass is seeded with d(total)/d(decoder output) directly.  The KL terms are seeded on the latent op.
  * Loss values stay on the device; nothing in here synchronises with the host.
"""
import numpy as np
import os

import torch
import torch.nn as nn

from . import ops, stack
from .fk_layer import ForwardKinematicsLayer, load_smpl24
from .skeleton import SkeletonConv, SkeletonPool, SkeletonUnpool, find_neighbor, get_edges


def _level_timesteps(args):
    """Per-level time lengths and encoder strides (seq_two_hier_sa_vae.py:76-90, 107-118)."""
    t, n = args['train_seq_len'], args['num_layers']
    steps, strides = [t], []
    for i in range(n):
        if t == 8:
            s = 1 if (i == 0 or i == n - 1) else 2
        elif t == 16:
            s = 1 if i == 0 else 2
        else:
            s = 2
        strides.append(s)
        steps.append(steps[-1] // s)
    return steps, strides


class Encoder(nn.Module):
    def __init__(self, args, topology):
        super(Encoder, self).__init__()
        self.topologies = [topology]
        self.latent_d = args['latent_d']
        self.shallow_latent_d = args['shallow_latent_d']
        self.channel_base = [6]
        self.channel_list = []
        self.edge_num = [len(topology)]
        self.pooling_list = []
        self.layers = nn.ModuleList()
        self.latent_enc_layers = nn.ModuleList()
        self.args = args
        self.convs = []
        self.pools = []
        n = args['num_layers']
        kernel_size = args['kernel_size']
        padding = (kernel_size - 1) // 2
        self.timestep_list, strides = _level_timesteps(args)
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
                                    joint_num=self.edge_num[i], kernel_size=kernel_size, stride=strides[i], padding=padding,
                                    padding_mode=args['padding_mode'], bias=True))
            self.convs.append(seq[-1])
            pool = SkeletonPool(edges=self.topologies[i], pooling_mode=args['skeleton_pool'],
                                channels_per_edge=out_channels // len(neighbor_list), last_pool=(i == n - 1))
            self.pools.append(pool)
            seq.append(pool)
            seq.append(nn.LeakyReLU(negative_slope=0.2))
            self.layers.append(nn.Sequential(*seq))
            lat = self.shallow_latent_d if i == 0 else self.latent_d
            self.latent_enc_layers.append(nn.Linear(self.channel_base[i + 1] * self.timestep_list[i + 1], lat * 2))
            self.topologies.append(pool.new_edges)
            self.pooling_list.append(pool.pooling_list)
            self.edge_num.append(len(self.topologies[-1]))

    def _identity_pool(self, i):
        return all(len(p) == 1 and p[0] == k for k, p in enumerate(self.pools[i].pooling_list))

    def conv_plans(self):
        """[(plan, weight)] of the main convs in forward order (for ops.prefetch_packs)."""
        return [(c.plan(lrelu=True) if self._identity_pool(i) else c.plan(), c.weight) for i, c in enumerate(self.convs)]

    def _level(self, i, x):
        seq = self.layers[i]
        for m in list(seq)[:-3]:              # extra convs (none in the shipped configs)
            x = m(x)
        conv, pool = self.convs[i], self.pools[i]
        if self._identity_pool(i):
            return conv.fused_forward(x, lrelu=True)          # identity pool (last level): conv + LeakyReLU epilogue
        return pool(conv(x), lrelu=True)                       # pool + LeakyReLU in one kernel

    def forward(self, input, offset=None, needed=None):
        """input [B, 24*6, T] -> (latent, [z_vec per level]); z_vec = [B, k_edges, 2*latent].

        ``needed``: optional set of level indices whose latent heads are required (others return None)."""
        z_vector_list = []
        levels = stack.encoder_forward(self, input, needed)  # all level outputs through the linked stack, or None
        for i in range(len(self.layers)):
            input = levels[i] if levels is not None else self._level(i, input)
            if needed is not None and i not in needed:
                z_vector_list.append(None)
                continue
            bs = input.shape[0]
            k_edges = input.shape[1] // self.channel_base[i + 1]
            head = self.latent_enc_layers[i]
            z_vector_list.append(ops.linear(input.view(bs, k_edges, -1), head.weight, head.bias))
        return input, z_vector_list


class Decoder(nn.Module):
    def __init__(self, args, enc: Encoder):
        super(Decoder, self).__init__()
        self.layers = nn.ModuleList()
        self.unpools = nn.ModuleList()
        self.latent_dec_layers = nn.ModuleList()
        self.latent_d = args['latent_d']
        self.shallow_latent_d = args['shallow_latent_d']
        self.args = args
        self.enc = enc
        self.convs = []
        self.hp = args
        self.upsample = []
        n = args['num_layers']
        self.timestep_list = list(reversed(_level_timesteps(args)[0]))
        kernel_size = args['kernel_size']
        padding = (kernel_size - 1) // 2
        for i in range(n):
            last = i == n - 1
            in_channels = enc.channel_list[n - i] * (2 if last else 1)
            out_channels = in_channels // (4 if last else 2)
            neighbor_list = find_neighbor(enc.topologies[n - i - 1], args['skeleton_dist'])
            bias = (i == 0 or last)
            lat = self.shallow_latent_d if last else self.latent_d
            self.latent_dec_layers.append(nn.Linear(lat, enc.channel_base[n - i] * self.timestep_list[i]))
            self.unpools.append(SkeletonUnpool(enc.pooling_list[n - i - 1], in_channels // len(neighbor_list)))
            if args['train_seq_len'] == 8:
                up = (not last) and i != 0
            elif args['train_seq_len'] == 16:
                up = not last
            else:
                up = True
            self.upsample.append(up)
            seq = []
            if up:
                seq.append(nn.Upsample(scale_factor=2, mode=args['upsampling'], align_corners=False))
            seq.append(self.unpools[-1])
            for _ in range(args['extra_conv']):
                seq.append(SkeletonConv(neighbor_list, in_channels=in_channels, out_channels=in_channels,
                                        joint_num=enc.edge_num[n - i - 1], kernel_size=kernel_size, stride=1, padding=padding,
                                        padding_mode=args['padding_mode'], bias=bias))
            seq.append(SkeletonConv(neighbor_list, in_channels=in_channels, out_channels=out_channels,
                                    joint_num=enc.edge_num[n - i - 1], kernel_size=kernel_size, stride=1, padding=padding,
                                    padding_mode=args['padding_mode'], bias=bias))
            self.convs.append(seq[-1])
            if not last:
                seq.append(nn.LeakyReLU(negative_slope=0.2))
            self.layers.append(nn.Sequential(*seq))
        if args.get('upsampling', 'linear') != 'linear':
            raise NotImplementedError("only upsampling='linear' (all shipped configs) is built")

    def _fused_kwargs(self, i):
        unpool = self.unpools[i]
        return dict(upsample=self.upsample[i], unpool_src=unpool.src, src_joints=unpool.input_edge_num,
                    lrelu=(i != self.hp['num_layers'] - 1))

    def conv_plans(self):
        if self.hp['extra_conv']:
            return []
        last = stack._last_plans.get(self)                  # the linked stack's plans once it has run (stack.decoder_forward)
        if last is not None and stack._enabled and ops._conv_impl != ops.IMPL_SIMT:
            return [(p, c.weight) for p, c in zip(last, self.convs)]
        return [(c.plan(**self._fused_kwargs(i)), c.weight) for i, c in enumerate(self.convs)]

    def _level(self, i, x):
        n = self.hp['num_layers']
        conv, unpool = self.convs[i], self.unpools[i]
        if self.hp['extra_conv']:
            # unfused fallback ordering for extra convs: upsample -> unpool -> convs
            if self.upsample[i]:
                x = ops.upsample2_linear(x)
            x = unpool(x)
            for m in self.layers[i]:
                if isinstance(m, SkeletonConv) and m is not conv:
                    x = m(x)
            return conv.fused_forward(x, lrelu=(i != n - 1))
        return conv.fused_forward(x, **self._fused_kwargs(i))

    def forward(self, z_vec_list, offset=None):
        """z_vec_list: shallow -> deep, each [B, k_edges, latent]; entries 1..n-2 may be None (they are dead inputs)."""
        n = len(z_vec_list)

        def feats(z_idx):
            z = z_vec_list[n - z_idx - 1]
            head = self.latent_dec_layers[z_idx]
            f = ops.linear(z, head.weight, head.bias)
            return f.view(z.size(0), -1, self.timestep_list[z_idx])

        x = feats(0)
        nl = self.hp['num_layers']
        if nl > 1 and len(self.layers) == nl:
            out = stack.decoder_forward(self, x, feats(nl - 1))
            if out is not None:
                return out
        for i in range(len(self.layers)):
            if i == self.hp['num_layers'] - 1 and i != 0:
                bs, _, t = x.size()
                k_edges = self.enc.edge_num[self.hp['num_layers'] - i]
                x = torch.cat((x.view(bs, k_edges, -1, t), feats(i).view(bs, k_edges, -1, t)), dim=2).view(bs, -1, t)
            x = self._level(i, x)
        return x


class TwoHierSAVAEModel(nn.Module):
    def __init__(self, hp, parent_json=None, device=None):
        super(TwoHierSAVAEModel, self).__init__()
        self.latent_d = hp['latent_d']
        self.shallow_latent_d = hp['shallow_latent_d']
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
        self.dec = Decoder(hp, self.enc)
        self.iteration_interval = hp['iteration_interval']
        mean_std = mean_std.copy()
        mean_std[1, mean_std[1, :] == 0] = 1.0
        self.mean_vals = torch.from_numpy(mean_std[0, :]).float()[None, :].to(dev)
        self.std_vals = torch.from_numpy(mean_std[1, :]).float()[None, :].to(dev)
        self._parents = parents
        self._const = {}
        self.mid_backward = None      # optional callback: decoder gradients are final (set by Trainer for the split optimiser step)

    def get_hier_level(self, step):
        return 1 if step < self.iteration_interval else 4

    # ------------------------------------------------------------------ helpers
    def _scalar(self, value, device):
        key = (float(value), str(device))
        t = self._const.get(key)
        if t is None:
            t = torch.full((), float(value), device=device, dtype=torch.float32)
            self._const[key] = t
        return t

    def _loss_buffers(self, device):
        key = ("loss", str(device))
        if key not in self._const:
            self._const[key] = (torch.zeros(8, device=device, dtype=torch.float32), torch.zeros(8, device=device, dtype=torch.float32))
        return self._const[key]

    def _draw_eps(self, z_vec_shapes, device, eps_list):
        """Four randn draws in level order (shapes [B*k_edges, d]) -- the reference's RNG consumption order
        (seq_two_hier_sa_vae.py:419-423).  ``eps_list`` injects them instead (parity tests)."""
        if eps_list is not None:
            return [e.to(device=device, dtype=torch.float32).contiguous() if e is not None else None for e in eps_list]
        return [torch.randn(rows, d, device=device, dtype=torch.float32) for rows, d in z_vec_shapes]

    def _to_device(self, t):
        dev = self.mean_vals.device
        return t.to(device=dev, dtype=torch.float32, non_blocking=True)

    # ------------------------------------------------------------------ training / validation step
    def forward(self, data, hp, iterations, multigpus=False, validation_flag=False, eps_list=None):
        seq_rot_6d, seq_rot_mat = data[0], data[1]
        seq_rot_6d = self._to_device(seq_rot_6d).contiguous()      # bs X T X (24*6)
        seq_rot_mat = self._to_device(seq_rot_mat).contiguous()    # bs X T X (24*3*3)
        bs, timesteps, _ = seq_rot_6d.size()
        dev = seq_rot_6d.device
        n = hp['num_layers']
        detach_shallow = iterations < hp['iteration_interval']

        k_edges = [len(p) for p in self.enc.pooling_list]
        lat = [self.shallow_latent_d] + [self.latent_d] * (n - 1)
        eps, eps_ready = [None] * n, None
        draw = None                            # deferred draw: () -> (eps, ready event); called once, after the encoder is issued
        if hp['kl_w'] != 0:
            if eps_list is not None:
                eps = self._draw_eps(None, dev, eps_list)
            else:
                # The four N(0,1) draws depend on nothing: they run on a side stream that forks HERE, off the encoder's critical
                # path -- but they are ISSUED only after the encoder's kernels: a replayed CUDA graph hands its first nodes to
                # the GPU one after the other (~2 us each, tools/timeline.py), and four generator kernels ahead of the first
                # conv delayed it by as much.
                fork = torch.cuda.Event()          # recorded below, after the step's first kernel: a dependency-free node at the
                                                   # head of a replayed graph is dispatched ahead of the forward pass's kernels

                def draw(_main=torch.cuda.current_stream()):
                    side = ops._eps_stream()
                    side.wait_event(fork)
                    with torch.cuda.stream(side):
                        e = self._draw_eps([(bs * k_edges[i], lat[i]) for i in range(n)], dev, None)
                        ready = torch.cuda.Event()
                        ready.record()
                    for t in e:
                        t.record_stream(_main)
                    return e, ready
                if os.environ.get("HMVAE_LATE_EPS", "1") == "0":
                    fork.record()
                    (eps, eps_ready), draw = draw(), None
        late = {"eps": eps, "ready": eps_ready, "draw": draw}
        ops.prefetch_packs(self.enc.conv_plans() + self.dec.conv_plans())   # tf32 weight copies, on the side stream
        x = ops.transpose_ct(seq_rot_6d)                            # bs X (24*6) X T   (input is [B, T, C])
        if late["draw"] is not None:
            fork.record()
        # persistent accumulators: [sum sq 6d, sum sq rot, sum sq pos, unused, KL sum shallow, KL sum deep]; zeroed by finalize
        acc, res = self._loss_buffers(dev)
        split = self.mid_backward is not None and not validation_flag
        out = self._fused_bottleneck(x, late, acc, hp, detach_shallow, split) if n > 1 else None
        fused = out is not None
        if fused:
            out, enc_outs, enc_cut = out
        else:
            _, z_vec_list = self.enc(x, needed={0, n - 1})
            eps, eps_ready = self._eps_now(late)
            if eps_ready is not None:
                torch.cuda.current_stream().wait_event(eps_ready)
            z_list = [None] * n
            levels = (0, n - 1) if n > 1 else (0,)
            for zi in levels:
                dist = z_vec_list[zi]
                deep = zi == n - 1
                kl_w = hp['kl_w'] if deep else hp['shallow_kl_w']
                slot = acc[5:6] if deep else acc[4:5]
                if (not deep) and detach_shallow:
                    with torch.no_grad():
                        z = ops.latent_fused(dist.detach(), eps[zi], lat[zi], 0.0, slot)
                else:
                    z = ops.latent_fused(dist, eps[zi], lat[zi], kl_w / (bs * k_edges[zi]), slot)
                z_list[zi] = z.view(bs, k_edges[zi], -1)

            # Split backward (when a mid-backward callback is installed): the decoder runs on detached latents, so that its
            # backward finishes -- and its optimiser / collective share can start -- before the encoder's backward begins.
            z_enc = list(z_list)
            if split:
                z_list = [z.detach().requires_grad_(z.requires_grad) if z is not None else None for z in z_list]
            out = self.dec(z_list)                                      # bs X (24*6) X T
        fk_off = self.fk_layer.positions[0].contiguous()
        dx6 = ops.recon_fwdbwd(out.detach(), True, seq_rot_6d, seq_rot_mat, fk_off, self._parents, hp['rec_6d_w'],
                               hp['rec_rot_w'], hp['rec_pose_w'], acc, want_grad=not validation_flag)
        nf = float(bs * timesteps)
        j = self.n_joints
        kl_deep_w = hp['kl_w'] if n > 1 else 0.0
        scale = [1.0 / (nf * 6 * j), 1.0 / (nf * 9 * j), 1.0 / (nf * 3 * j), 0.0, 1.0 / (bs * k_edges[0]), 1.0 / (bs * k_edges[n - 1])]
        w_all = [hp['rec_6d_w'], hp['rec_rot_w'], hp['rec_pose_w'], 0.0, hp['shallow_kl_w'] if n > 1 else hp['kl_w'], kl_deep_w]
        w_kl = [0.0, 0.0, 0.0, 0.0, w_all[4], w_all[5]]
        ops.loss_finalize(acc, res, scale, w_all, w_kl)              # one kernel; res = [l6, lrot, lpos, -, kl0, kl3, total, kl]
        l_rec_6d, l_rec_rot_mat, l_rec_pose = res[0], res[1], res[2]
        zero = res[3:4]
        l_kl_list = [res[4]] + [zero for _ in range(max(n - 2, 0))] + ([res[5]] if n > 1 else [])
        l_total, l_kl = res[6], res[7]

        if not validation_flag:
            with ops.wgrad_overlap():
                out.backward(dx6)
                if split and fused:
                    self.mid_backward()
                    pairs = [(a, c.grad) for a, c in zip(enc_outs, enc_cut) if c.grad is not None]
                    if pairs:
                        torch.autograd.backward([a for a, _ in pairs], [g for _, g in pairs])
                elif split:
                    self.mid_backward()
                    pairs = [(ze, zd.grad) for ze, zd in zip(z_enc, z_list) if ze is not None and ze.requires_grad and zd.grad is not None]
                    if pairs:
                        torch.autograd.backward([a for a, _ in pairs], [g for _, g in pairs])

        return l_total, l_kl, l_rec_6d, l_rec_rot_mat, l_rec_pose, zero, zero, zero, zero, l_kl_list

    @staticmethod
    def _eps_now(late):
        """Issues the deferred N(0,1) draws (once)."""
        if late["draw"] is not None:
            (late["eps"], late["ready"]), late["draw"] = late["draw"](), None
        return late["eps"], late["ready"]

    def _fused_bottleneck(self, x, late, acc, hp, detach_shallow, split):
        """Encoder stack -> both latent heads + reparametrisation + KL + both decoder heads in one kernel -> decoder stack (the
        linked tensor-core path, stack.py).  Returns (decoder output, encoder outputs fed to the heads, their detached twins when
        ``split``) or None when the linked path cannot run this geometry (caller takes the per-layer path)."""
        from . import stack
        n = hp['num_layers']
        enc, dec = self.enc, self.dec
        if len(dec.layers) != n or hp['extra_conv']:
            return None
        outs = stack.encoder_forward(enc, x, {0, n - 1})
        if outs is None:
            return None
        bs = x.shape[0]
        k_edges = [len(p) for p in enc.pooling_list]
        eps, eps_ready = self._eps_now(late)
        if eps_ready is not None:
            torch.cuda.current_stream().wait_event(eps_ready)
        srcs = [outs[0], outs[n - 1]]                       # shallow (level 0) and deep (level n-1) features
        cut = srcs
        if split:                                            # decoder + heads backward first, then the optimiser's decoder share
            cut = [t.detach().requires_grad_(True) for t in srcs]
        d_sh, d_dp = self.shallow_latent_d, self.latent_d
        metas = [dict(d=d_sh, kl_scale=hp['shallow_kl_w'] / (bs * k_edges[0]), kl_acc=acc[4:5], detach=bool(detach_shallow)),
                 dict(d=d_dp, kl_scale=hp['kl_w'] / (bs * k_edges[n - 1]), kl_acc=acc[5:6], detach=False)]
        he0, he3 = enc.latent_enc_layers[0], enc.latent_enc_layers[n - 1]
        hd_sh, hd_dp = dec.latent_dec_layers[n - 1], dec.latent_dec_layers[0]      # decoder heads: index 0 takes the deep latent
        tensors = [cut[0].view(bs, k_edges[0], -1), he0.weight, he0.bias, hd_sh.weight, hd_sh.bias, eps[0],
                   cut[1].view(bs, k_edges[n - 1], -1), he3.weight, he3.bias, hd_dp.weight, hd_dp.bias, eps[n - 1]]
        feats, _ = ops.latent_heads(metas, tensors)
        feat_sh = feats[0].view(bs, -1, dec.timestep_list[n - 1])
        feat_dp = feats[1].view(bs, -1, dec.timestep_list[0])
        out = stack.decoder_forward(dec, feat_dp, feat_sh)
        if out is None:
            return None
        return out, srcs, cut

    # ------------------------------------------------------------------ reference helpers kept for callers
    def reparametrize(self, pred_mean, pred_logvar):
        dist = torch.cat([pred_mean, pred_logvar], dim=1)
        z, _ = ops.latent_sample_kl(dist, torch.randn_like(pred_mean), pred_mean.shape[1])
        return z

    def kl_loss(self, logvar, mu):
        _, kl = ops.latent_sample_kl(torch.cat([mu, logvar], dim=1), None, mu.shape[1])
        return kl / mu.shape[0]

    def l2_criterion(self, pred, gt):
        return ops.l2_criterion(pred, gt)

    def _decode(self, z_list, adjust_root_rot_flag=False, relative_root_rot=None):
        """seq_two_hier_sa_vae.py:436-474: decoder -> [B,T,24,6] -> rotation matrices -> FK positions."""
        result = self.dec(z_list)                                   # bs X (24*out_dim) X T
        bs = result.size(0)
        decoder_out = ops.transpose_ct(result).view(bs * self.max_timesteps, self.n_joints, -1)
        cont6d_rep = decoder_out[:, :, :self.output_dim]
        out_rotation_matrix = ops.rot6d_to_rotmat(cont6d_rep)
        if adjust_root_rot_flag:
            out_rotation_matrix = out_rotation_matrix.view(bs, self.max_timesteps, 24, 3, 3)
            if relative_root_rot is not None:
                out_rotation_matrix = out_rotation_matrix.clone()
                out_rotation_matrix[:, :, 0, :, :] = torch.matmul(relative_root_rot, out_rotation_matrix[:, :, 0, :, :])
            else:
                out_rotation_matrix, relative_root_rot = self.adjust_root_rot(out_rotation_matrix)
            out_rotation_matrix = out_rotation_matrix.reshape(bs * self.max_timesteps, 24, 3, 3)
        out_pose_pos = self.fk_layer(out_rotation_matrix)
        out_cont6d = cont6d_rep.view(bs, self.max_timesteps, self.n_joints, -1)
        out_rotation_matrix = out_rotation_matrix.view(bs, self.max_timesteps, self.n_joints, 3, 3)
        out_pose_pos = out_pose_pos.view(bs, self.max_timesteps, self.n_joints, 3)
        return out_cont6d, out_rotation_matrix, out_pose_pos, None, None, None, None

    def _decode_w_given_decoder(self, z_list, curr_decoder):
        """seq_two_hier_sa_vae.py:501-529."""
        from .latent_opt import decode_w_given_decoder
        return decode_w_given_decoder(self, z_list, curr_decoder)

    def l2_masked_criterion(self, pred, gt, mask):
        """seq_two_hier_sa_vae.py:717-735 -> (mean over all elements of (pred-gt)^2 * mask, per-frame mean [bs, T])."""
        from .latent_opt import l2_masked_criterion
        return l2_masked_criterion(pred, gt, mask)

    def optimize_latent(self, target_cont6d, target_rotmat, target_mask, hp=None, z_vec_list=None, prev_epochs=50, opt_it=None,
                        cuda_graph=True):
        """The latent-space optimisation loop shared by final_long_seq_try_interpolation (seq_two_hier_sa_vae.py:1356-1429,
        prev_epochs = 50) and final_motion_completion_long_seq (:1698-1757, prev_epochs = 100) for one window of
        ``max_timesteps`` frames: targets bs X T X 24 X 6 / bs X T X 24 X 3 X 3, mask bs X T X 24.  See latent_opt.py."""
        from .latent_opt import optimize_latent
        return optimize_latent(self, target_cont6d, target_rotmat, target_mask, self.hp if hp is None else hp,
                               z_vec_list=z_vec_list, prev_epochs=prev_epochs, opt_it=opt_it, cuda_graph=cuda_graph)

    def adjust_root_rot(self, ori_seq_data):
        """seq_two_hier_sa_vae.py:531-551: rotate every frame's root so that frame 0 faces the identity."""
        bs, timesteps = ori_seq_data.shape[:2]
        relative_rot = ori_seq_data[:, 0, 0].transpose(1, 2)[:, None].repeat(1, timesteps, 1, 1)
        converted = torch.matmul(relative_rot, ori_seq_data[:, :, 0])
        dest = ori_seq_data.clone()
        dest[:, :, 0] = converted
        return dest, relative_rot

    def de_standardize(self, output_data, start_idx, end_idx):
        if output_data.dim() == 2:
            return self.mean_vals[:, start_idx:end_idx] + self.std_vals[:, start_idx:end_idx] * output_data
        return self.mean_vals[None][:, :, start_idx:end_idx] + self.std_vals[None][:, :, start_idx:end_idx] * output_data

    def test(self, data, hp, iterations, gen_seq_len=None, sampled_z_list=None):
        """seq_two_hier_sa_vae.py:560-639: 1 encoder pass + 2 decoder passes + 3 FK, no grad.
        Returns (gt, mean-decoded, sampled-decoded) joint positions as T X bs X 24 X 3."""
        was_training = self.training
        self.eval()
        with torch.no_grad():
            seq_rot_6d = self._to_device(data[0]).contiguous()
            seq_rot_mat = self._to_device(data[1]).contiguous()
            bs, timesteps, _ = seq_rot_6d.size()
            relative_rot = None
            if hp['random_root_rot_flag']:
                rm, relative_rot = self.adjust_root_rot(seq_rot_mat.view(bs, timesteps, 24, 3, 3))
                seq_rot_mat = rm.reshape(bs, timesteps, -1)
            seq_rot_pos = self.fk_layer(seq_rot_mat.view(bs * timesteps, self.n_joints, 3, 3))
            gt_seq_res = seq_rot_pos.view(bs, timesteps, 24, 3).transpose(0, 1).contiguous()
            nl = hp['num_layers']
            # only the shallow and the deep latent reach the decoder (reference :278-288): the two middle heads are dead work
            _, z_vec_list = self.enc(ops.transpose_ct(seq_rot_6d), needed={0, nl - 1})
            n = len(z_vec_list)
            mean_z_list, sampled = [], []
            for zi, dist in enumerate(z_vec_list):
                if dist is None:
                    if sampled_z_list is None:                       # keep the reference's RNG consumption: one draw per level
                        torch.randn(bs, len(self.enc.pooling_list[zi]), self.latent_d, device=seq_rot_6d.device)
                    mean_z_list.append(None)
                    sampled.append(None)
                    continue
                d = self.shallow_latent_d if zi == 0 else self.latent_d
                mean_z = dist[:, :, :d].contiguous()
                mean_z_list.append(mean_z)
                sampled.append(torch.randn_like(mean_z) if sampled_z_list is None else sampled_z_list[zi].to(mean_z.device))
            if hp['random_root_rot_flag']:
                mean_pos = self._decode(mean_z_list, adjust_root_rot_flag=True, relative_root_rot=relative_rot)[2]
                samp_pos = self._decode(sampled, adjust_root_rot_flag=True)[2]
            else:
                mean_pos = self._decode(mean_z_list)[2]
                samp_pos = self._decode(sampled)[2]
        self.train(was_training)
        return gt_seq_res, mean_pos.transpose(0, 1), samp_pos.transpose(0, 1), None

    def gen_seq(self, data, hp, iterations):
        return self.test(data, hp, iterations, hp['max_input_timesteps'])

    def aa2matrot(self, pose):
        """seq_two_hier_sa_vae.py:644-654: Nx1xJx3 axis-angle -> NxJx3x3."""
        batch_size = pose.size(0)
        m = ops.angle_axis_to_rotation_matrix(pose.reshape(-1, 3).float())[:, :3, :3].contiguous()
        return m.view(batch_size, self.n_joints, 3, 3)

    def aa2others(self, aa_data):
        """seq_two_hier_sa_vae.py:656-675: [bs,T,72] axis-angle -> (6D, rotmat, FK positions)."""
        bs, timesteps, _ = aa_data.size()
        rot = self.aa2matrot(aa_data.view(bs * timesteps, self.n_joints, 3)[:, None])
        cont6d = torch.stack((rot[:, :, :, 0], rot[:, :, :, 1]), dim=-2).view(rot.size(0), rot.size(1), 6)
        pose_pos = self.fk_layer(rot)
        return cont6d.view(bs, timesteps, -1), rot.view(bs, timesteps, -1), pose_pos.contiguous().view(bs, timesteps, -1)
