"""Oracle (test infrastructure): torch fp32 restatement of the hm-vae hot path.

Device-agnostic (runs on CPU for the parity tests and the CPU baseline).  Each function
cites the reference lines it restates.  PINNED: ``tests/golden/*.npz`` were produced by
``oracle/make_golden.py`` from the real reference modules imported from /root/reference;
``tests/test_oracle_golden.py`` checks this file against them.  The one exception is
``angle_axis_to_rotation_matrix`` (torchgeometry is not vendored by the reference and not
installed): PARITY UNPINNED against torchgeometry itself; pinned only against scipy's
Rodrigues away from theta ~ 0.

Never imported by the product package.
"""
import math

import torch
import torch.nn.functional as F

from . import topology as topo

# --------------------------------------------------------------------------------------
# elementary ops
# --------------------------------------------------------------------------------------


def skeleton_conv(x, weight, mask, bias, stride, padding, padding_mode="reflect"):
    """skeleton.py:95-105 -- conv1d(pad(x), W*mask, b, stride)."""
    mode = {"zeros": "constant", "constant": "constant", "reflection": "reflect", "reflect": "reflect"}[padding_mode]
    xp = F.pad(x, (padding, padding), mode=mode)
    return F.conv1d(xp, weight * mask, bias, stride, 0, 1, 1)


def conv_mask(neigh, cin, cout, ksize):
    """skeleton.py:34-39, 58-61."""
    j = len(neigh)
    ci, co = cin // j, cout // j
    m = torch.zeros(cout, cin, ksize)
    for jo, nb in enumerate(neigh):
        for k in nb:
            m[jo * co:(jo + 1) * co, k * ci:(k + 1) * ci, :] = 1
    return m


def skeleton_pool(x, pooling_list, c):
    """skeleton.py:219-231 -- mean over the member edges' channel blocks."""
    outs = []
    for members in pooling_list:
        acc = None
        for j in members:
            blk = x[:, j * c:(j + 1) * c, :] * (1.0 / len(members))
            acc = blk if acc is None else acc + blk
        outs.append(acc)
    return torch.cat(outs, dim=1)


def skeleton_unpool(x, pooling_list, c):
    """skeleton.py:248-261 -- copy each pooled edge's channels to its members."""
    n_out = sum(len(m) for m in pooling_list)
    src = [0] * n_out
    for i, members in enumerate(pooling_list):
        for j in members:
            src[j] = i
    return torch.cat([x[:, s * c:(s + 1) * c, :] for s in src], dim=1)


def pool_weight(pooling_list, c, n_edges):
    w = torch.zeros(len(pooling_list) * c, n_edges * c)
    for i, members in enumerate(pooling_list):
        for j in members:
            for ch in range(c):
                w[i * c + ch, j * c + ch] = 1.0 / len(members)
    return w


def unpool_weight(pooling_list, c):
    n_out = sum(len(m) for m in pooling_list)
    w = torch.zeros(n_out * c, len(pooling_list) * c)
    for i, members in enumerate(pooling_list):
        for j in members:
            for ch in range(c):
                w[j * c + ch, i * c + ch] = 1
    return w


def upsample2_linear(x):
    """nn.Upsample(scale_factor=2, mode='linear', align_corners=False) -- seq_two_hier_sa_vae.py:233-240."""
    t = x.shape[-1]
    idx = torch.arange(t, device=x.device)
    left = x[..., (idx - 1).clamp(min=0)]
    right = x[..., (idx + 1).clamp(max=t - 1)]
    even = 0.75 * x + 0.25 * left
    odd = 0.75 * x + 0.25 * right
    return torch.stack([even, odd], dim=-1).reshape(*x.shape[:-1], 2 * t)


def rot6d_to_rotmat(x6):
    """my_tools.py:19-39 -- columns [x, y, z]; normalize = v / max(|v|, 1e-6)."""
    a, b = x6[..., 0:3], x6[..., 3:6]
    x = a / a.norm(dim=-1, keepdim=True).clamp_min(1e-6)
    z = torch.cross(x, b, dim=-1)
    z = z / z.norm(dim=-1, keepdim=True).clamp_min(1e-6)
    y = torch.cross(z, x, dim=-1)
    return torch.stack([x, y, z], dim=-1)


def rotmat_to_rot6d(r):
    """Inlined everywhere in the reference, e.g. seq_two_hier_sa_vae.py:666-667."""
    return torch.stack([r[..., 0], r[..., 1]], dim=-2).reshape(*r.shape[:-2], 6)


def forward_kinematics(rot, parents, offsets):
    """fk_layer.py:47-93.  rot: [N,J,3,3] or [N,J,6]; offsets [J,3] or [N,J,3] -> [N,J,3]."""
    if rot.shape[-1] == 6:
        rot = rot6d_to_rotmat(rot)
    n, j = rot.shape[0], rot.shape[1]
    if offsets.dim() == 2:
        offsets = offsets.unsqueeze(0).expand(n, j, 3)
    g_rot = [None] * j
    g_pos = [None] * j
    g_rot[0] = rot[:, 0]
    g_pos[0] = offsets[:, 0]
    for i in range(1, j):
        p = int(parents[i])
        g_pos[i] = torch.einsum("nab,nb->na", g_rot[p], offsets[:, i]) + g_pos[p]
        g_rot[i] = torch.matmul(g_rot[p], rot[:, i])
    return torch.stack(g_pos, dim=1)


def angle_axis_to_rotation_matrix(aa):
    """torchgeometry 0.1.2 ``angle_axis_to_rotation_matrix`` ([N,3] -> [N,4,4]); PARITY UNPINNED.

    Rodrigues with theta = sqrt(aa.aa), axis = aa/(theta+1e-6); where theta^2 <= 1e-6 the
    first-order matrix [[1,-rz,ry],[rz,1,-rx],[-ry,rx,1]] is used instead.
    Call sites: seq_two_hier_sa_vae.py:650, trajectory_pred_model.py:451.
    """
    eps = 1e-6
    theta2 = (aa * aa).sum(dim=1)
    theta = torch.sqrt(theta2)
    w = aa / (theta + eps).unsqueeze(1)
    wx, wy, wz = w[:, 0], w[:, 1], w[:, 2]
    c, s = torch.cos(theta), torch.sin(theta)
    k1 = 1.0
    r_normal = torch.stack([
        c + wx * wx * (k1 - c), wx * wy * (k1 - c) - wz * s, wy * s + wx * wz * (k1 - c),
        wz * s + wx * wy * (k1 - c), c + wy * wy * (k1 - c), -wx * s + wy * wz * (k1 - c),
        -wy * s + wx * wz * (k1 - c), wx * s + wy * wz * (k1 - c), c + wz * wz * (k1 - c)], dim=1).view(-1, 3, 3)
    rx, ry, rz = aa[:, 0], aa[:, 1], aa[:, 2]
    one = torch.ones_like(rx)
    r_taylor = torch.stack([one, -rz, ry, rz, one, -rx, -ry, rx, one], dim=1).view(-1, 3, 3)
    use_normal = (theta2 > eps).view(-1, 1, 1).to(aa.dtype)
    r = use_normal * r_normal + (1 - use_normal) * r_taylor
    out = torch.eye(4, dtype=aa.dtype, device=aa.device).repeat(aa.shape[0], 1, 1)
    out[:, :3, :3] = r
    return out


def kl_loss(logvar, mu):
    """seq_two_hier_sa_vae.py:425-428."""
    return (-0.5 * torch.sum(1 + logvar - mu.pow(2) - logvar.exp(), dim=1)).mean()


def l2_criterion(pred, gt):
    """seq_two_hier_sa_vae.py:430-434."""
    assert pred.size() == gt.size()
    return ((pred - gt) ** 2).mean()


def l2_masked_criterion(pred, gt, mask):
    """seq_two_hier_sa_vae.py:717-735.  pred/gt: bs X T X 24 X {6 | 3 X 3 | 3}; mask: bs X T X 24 -> (loss, [bs, T])."""
    assert pred.size() == gt.size()
    m = mask[:, :, :, None] if pred.dim() == 4 else mask[:, :, :, None, None]
    loss = (pred - gt) ** 2 * m
    return loss.mean(), loss.reshape(loss.shape[0], loss.shape[1], -1).mean(dim=-1)


# --------------------------------------------------------------------------------------
# parameter construction with the reference's RNG consumption order
# --------------------------------------------------------------------------------------


def _init_skeleton_conv(neigh, cin, cout, ksize, bias):
    """skeleton.py:70-89 (per output joint: kaiming_uniform(a=sqrt5) on the unmasked block, then bias)."""
    j = len(neigh)
    ci, co = cin // j, cout // j
    w = torch.zeros(cout, cin, ksize)
    b = torch.zeros(cout) if bias else None
    for jo, nb in enumerate(neigh):
        cols = [k * ci + c for k in nb for c in range(ci)]
        tmp = torch.zeros(co, len(cols), ksize)
        torch.nn.init.kaiming_uniform_(tmp, a=math.sqrt(5))
        w[jo * co:(jo + 1) * co, cols, :] = tmp
        if bias:
            bound = 1.0 / math.sqrt(len(cols) * ksize)
            tb = torch.zeros(co)
            torch.nn.init.uniform_(tb, -bound, bound)
            b[jo * co:(jo + 1) * co] = tb
    return w, b


def _init_linear(fin, fout):
    lin = torch.nn.Linear(fin, fout)  # default init consumes RNG exactly as the reference's nn.Linear does
    return lin.weight.detach().clone(), lin.bias.detach().clone()


def _kaiming_linear_(w, b):
    """trainer_motion_vae.py:264-283 (weights_init('kaiming') touches nn.Linear only)."""
    torch.nn.init.kaiming_normal_(w, a=0, mode="fan_in")
    b.zero_()


def hmvae_timesteps(hp):
    """seq_two_hier_sa_vae.py:76-90 -- per-level time lengths and encoder strides."""
    t, n = hp["train_seq_len"], hp["num_layers"]
    ts, strides = [t], []
    for i in range(n):
        if t == 8:
            s = 1 if (i == 0 or i == n - 1) else 2
        elif t == 16:
            s = 1 if i == 0 else 2
        else:
            s = 2
        strides.append(s)
        ts.append(ts[-1] // s)
    return ts, strides


class HMVAEOracle:
    """Functional restatement of Encoder/Decoder/TwoHierSAVAEModel (seq_two_hier_sa_vae.py:53-474).

    ``params`` uses the reference's state_dict key names.  ``init(seed)`` reproduces the reference
    construction order (Encoder then Decoder, then Trainer's ``apply(weights_init)`` which re-draws
    the encoder's nn.Linear weights a second time through ``dec.enc``).
    """

    def __init__(self, hp, parents=topo.SMPL24_PARENTS, offsets=None):
        self.hp = hp
        self.parents = list(parents)
        self.offsets = offsets
        self.n = hp["num_layers"]
        self.levels = topo.hierarchy(parents, self.n, hp["skeleton_dist"])
        self.edge_num = [len(l["edges"]) for l in self.levels] + [len(self.levels[-1]["new_edges"])]
        self.ts, self.strides = hmvae_timesteps(hp)
        self.ksize = hp["kernel_size"]
        self.pad = (self.ksize - 1) // 2
        self.cbase = [6 * 2 ** i for i in range(self.n + 1)]
        self.channel_list = [self.cbase[0] * self.edge_num[0]] + [self.cbase[i + 1] * self.edge_num[i] for i in range(self.n)]
        self.dts = list(reversed(self.ts))
        self.up = []
        for i in range(self.n):
            if hp["train_seq_len"] == 8:
                self.up.append(i != self.n - 1 and i != 0)
            elif hp["train_seq_len"] == 16:
                self.up.append(i != self.n - 1)
            else:
                self.up.append(True)
        self.params = {}
        self.masks = {}

    # ---- shapes -------------------------------------------------------------------
    def enc_conv_shape(self, i):
        return self.cbase[i] * self.edge_num[i], self.cbase[i + 1] * self.edge_num[i]

    def dec_conv_shape(self, i):
        n = self.n
        cin = self.channel_list[n - i] * (2 if i == n - 1 else 1)
        cout = cin // 4 if i == n - 1 else cin // 2
        return cin, cout

    def dec_conv_key(self, i):
        return "dec.layers.%d.%d" % (i, 2 if self.up[i] else 1)

    # ---- init ---------------------------------------------------------------------
    def init(self, seed=0):
        hp, n, p = self.hp, self.n, {}
        torch.manual_seed(seed)
        for i in range(n):
            cin, cout = self.enc_conv_shape(i)
            w, b = _init_skeleton_conv(self.levels[i]["neighbours"], cin, cout, self.ksize, True)
            p["enc.layers.%d.0.weight" % i], p["enc.layers.%d.0.bias" % i] = w, b
            lat = hp["shallow_latent_d"] if i == 0 else hp["latent_d"]
            p["enc.latent_enc_layers.%d.weight" % i], p["enc.latent_enc_layers.%d.bias" % i] = \
                _init_linear(self.cbase[i + 1] * self.ts[i + 1], 2 * lat)
        for i in range(n):
            cin, cout = self.dec_conv_shape(i)
            lat = hp["shallow_latent_d"] if i == n - 1 else hp["latent_d"]
            p["dec.latent_dec_layers.%d.weight" % i], p["dec.latent_dec_layers.%d.bias" % i] = \
                _init_linear(lat, self.cbase[n - i] * self.dts[i])
            bias = (i == 0 or i == n - 1)
            w, b = _init_skeleton_conv(self.levels[n - 1 - i]["neighbours"], cin, cout, self.ksize, bias)
            p[self.dec_conv_key(i) + ".weight"] = w
            if bias:
                p[self.dec_conv_key(i) + ".bias"] = b
        # Trainer.apply(weights_init('kaiming')): enc linears, dec linears, then enc linears again (dec.enc)
        for i in range(n):
            _kaiming_linear_(p["enc.latent_enc_layers.%d.weight" % i], p["enc.latent_enc_layers.%d.bias" % i])
        for i in range(n):
            _kaiming_linear_(p["dec.latent_dec_layers.%d.weight" % i], p["dec.latent_dec_layers.%d.bias" % i])
        for i in range(n):
            _kaiming_linear_(p["enc.latent_enc_layers.%d.weight" % i], p["enc.latent_enc_layers.%d.bias" % i])
        self.set_params(p)
        return self

    def set_params(self, p, device=None):
        self.params = {k: v.detach().clone().to(device or v.device).requires_grad_(True) for k, v in p.items()
                       if not k.endswith("mask")}
        dev = next(iter(self.params.values())).device
        for i in range(self.n):
            cin, cout = self.enc_conv_shape(i)
            self.masks["enc.layers.%d.0" % i] = conv_mask(self.levels[i]["neighbours"], cin, cout, self.ksize).to(dev)
            cin, cout = self.dec_conv_shape(i)
            self.masks[self.dec_conv_key(i)] = conv_mask(self.levels[self.n - 1 - i]["neighbours"], cin, cout, self.ksize).to(dev)
        if self.offsets is not None:
            self.offsets = self.offsets.to(dev)

    def trainable(self):
        return self.params

    # ---- forward ------------------------------------------------------------------
    def encode(self, x):
        """Encoder.forward, seq_two_hier_sa_vae.py:142-167.  x: [B, 144, T]."""
        p, hp = self.params, self.hp
        zs = []
        for i in range(self.n):
            key = "enc.layers.%d.0" % i
            x = skeleton_conv(x, p[key + ".weight"], self.masks[key], p[key + ".bias"], self.strides[i], self.pad,
                              hp["padding_mode"])
            x = skeleton_pool(x, self.levels[i]["pooling_list"], self.cbase[i + 1])
            x = F.leaky_relu(x, 0.2)
            k_edges = x.shape[1] // self.cbase[i + 1]
            feat = x.reshape(x.shape[0], k_edges, -1)
            zs.append(F.linear(feat, p["enc.latent_enc_layers.%d.weight" % i], p["enc.latent_enc_layers.%d.bias" % i]))
        return x, zs

    def decode_net(self, z_list):
        """Decoder.forward, seq_two_hier_sa_vae.py:260-294."""
        p, n = self.params, self.n
        feats = []
        for zi in range(n):
            z = z_list[n - 1 - zi]
            f = F.linear(z, p["dec.latent_dec_layers.%d.weight" % zi], p["dec.latent_dec_layers.%d.bias" % zi])
            feats.append(f.reshape(z.shape[0], -1, self.dts[zi]))
        x = None
        for i in range(n):
            if i == 0:
                x = feats[0]
            elif i == n - 1:
                bs, _, t = x.shape
                k_edges = self.edge_num[n - i]
                x = torch.cat([x.reshape(bs, k_edges, -1, t), feats[i].reshape(bs, k_edges, -1, t)], dim=2).reshape(bs, -1, t)
            if self.up[i]:
                x = upsample2_linear(x)
            cin, _ = self.dec_conv_shape(i)
            lvl = self.levels[n - 1 - i]
            x = skeleton_unpool(x, lvl["pooling_list"], cin // len(lvl["neighbours"]))
            key = self.dec_conv_key(i)
            x = skeleton_conv(x, p[key + ".weight"], self.masks[key], p.get(key + ".bias"), 1, self.pad,
                              self.hp["padding_mode"])
            if i != n - 1:
                x = F.leaky_relu(x, 0.2)
        return x

    def decode(self, z_list):
        """_decode, seq_two_hier_sa_vae.py:436-474."""
        out = self.decode_net(z_list)
        bs, t = out.shape[0], out.shape[2]
        x6 = out.transpose(1, 2).contiguous().view(bs * t, len(self.parents), -1)
        rot = rot6d_to_rotmat(x6)
        pos = forward_kinematics(rot, self.parents, self.offsets)
        return x6.view(bs, t, -1), rot.view(bs, t, -1), pos.view(bs, t, -1)

    def latents(self, zs, eps_list, iterations):
        """seq_two_hier_sa_vae.py:357-391."""
        hp = self.hp
        z_list, kls = [], []
        for zi, dist in enumerate(zs):
            d = hp["shallow_latent_d"] if zi == 0 else hp["latent_d"]
            bs, k_edges, _ = dist.shape
            mu = dist[:, :, :d].reshape(-1, d)
            lv = dist[:, :, d:].reshape(-1, d)
            z = eps_list[zi] * torch.exp(0.5 * lv) + mu if hp["kl_w"] != 0 else mu
            z = z.view(bs, k_edges, -1)
            if zi == len(zs) - 1:
                kl = kl_loss(lv, mu)
            elif zi == 0:
                if iterations < hp["iteration_interval"]:
                    kl = kl_loss(lv.detach(), mu.detach())
                    z = z.detach()
                else:
                    kl = kl_loss(lv, mu)
            else:
                kl = torch.zeros(1, device=dist.device)
            z_list.append(z)
            kls.append(kl)
        return z_list, kls

    def step(self, seq_rot_6d, seq_rot_mat, eps_list, iterations=0, backward=True):
        """TwoHierSAVAEModel.forward, seq_two_hier_sa_vae.py:335-417 (eps injected instead of randn_like)."""
        hp = self.hp
        bs, t, _ = seq_rot_6d.shape
        j = len(self.parents)
        with torch.no_grad():
            gt_pos = forward_kinematics(seq_rot_mat.view(bs * t, j, 3, 3), self.parents, self.offsets).view(bs, t, -1)
        x = seq_rot_6d.view(bs, t, -1).transpose(1, 2)
        _, zs = self.encode(x)
        z_list, kls = self.latents(zs, eps_list, iterations)
        x6, rot, pos = self.decode(z_list)
        l6d = l2_criterion(x6, seq_rot_6d)
        lrot = l2_criterion(rot, seq_rot_mat)
        lpos = l2_criterion(pos, gt_pos)
        total = hp["rec_6d_w"] * l6d + hp["rec_rot_w"] * lrot + hp["rec_pose_w"] * lpos + hp["kl_w"] * kls[-1] \
            + hp["shallow_kl_w"] * kls[0]
        if backward:
            total.reshape(()).backward()
        return dict(total=total.reshape(()), rec_6d=l6d, rec_rot=lrot, rec_pose=lpos,
                    kl_deep=kls[-1].reshape(()), kl_shallow=kls[0].reshape(()), x6=x6, rot=rot, pos=pos)

    def latent_optimise(self, z_init, target_6d, target_rot, mask, hp, prev_epochs=50, opt_it=None):
        """Inner loop of final_long_seq_try_interpolation (seq_two_hier_sa_vae.py:1356-1429) / final_motion_completion_long_seq
        (:1698-1757): opt_it iterations; iterations i <= prev_epochs step Adam(z_vec_list, lr=opt_lr), later ones step
        Adam(curr_decoder.parameters(), lr=opt_lr*0.001) when hp['optimize_decoder'] (curr_decoder = deepcopy(self.dec), which
        carries the ``dec.enc.*`` encoder copies: they get the regulariser's gradient and weight decay like every other parameter);
        each optimiser has its own StepLR(opt_step_size, opt_gamma) when opt_lr_policy == 'step' (:40-51).

        z_init: 4 tensors [bs, k_edges, d]; target_6d bs X T X 24 X 6; target_rot bs X T X 24 X 3 X 3; mask bs X T X 24.
        Uses (and, in the decoder phase, UPDATES a copy of) self.params.  Returns dict(losses [opt_it, 6] = (rec_6d, rec_rot,
        rec_pose, reg, reg_decoder, total), out_6d / out_rot_mat / out_pose_pos of the LAST iteration, z, decoder_params)."""
        from torch.optim import lr_scheduler

        opt_it = hp["opt_it"] if opt_it is None else opt_it
        bs, t = target_6d.shape[0], target_6d.shape[1]
        j = len(self.parents)
        with torch.no_grad():
            target_pos = forward_kinematics(target_rot.reshape(bs * t, j, 3, 3), self.parents, self.offsets).view(bs, t, j, 3)
        z = [torch.nn.Parameter(v.detach().clone()) for v in z_init]
        orig = {k: v.detach().clone() for k, v in self.params.items()}
        # curr_decoder.named_parameters(): dec.* and, through dec.enc, the encoder's (trainable ones)
        cur = {k: torch.nn.Parameter(v.detach().clone()) for k, v in self.params.items()}
        saved, self.params = self.params, cur
        optimize_decoder = bool(hp.get("optimize_decoder", False))

        def sched(opt):
            if hp.get("opt_lr_policy", "constant") == "step":
                return lr_scheduler.StepLR(opt, step_size=hp["opt_step_size"], gamma=hp["opt_gamma"])
            return None

        z_opt = torch.optim.Adam(z, lr=hp["opt_lr"], weight_decay=hp["weight_decay"])
        z_sched = sched(z_opt)
        if optimize_decoder:
            d_opt = torch.optim.Adam(list(cur.values()), lr=hp["opt_lr"] * 0.001, weight_decay=hp["weight_decay"])
            d_sched = sched(d_opt)
        hist = []
        try:
            for i in range(opt_it):
                x6, rot, pos = self.decode(z)
                x6, rot, pos = x6.view(bs, t, j, 6), rot.view(bs, t, j, 3, 3), pos.view(bs, t, j, 3)
                l6, _ = l2_masked_criterion(x6, target_6d, mask)
                lrot, _ = l2_masked_criterion(rot, target_rot, mask)
                lpos, _ = l2_masked_criterion(pos, target_pos, mask)
                l_reg = l2_criterion(z[0], torch.zeros_like(z[0])) + l2_criterion(z[3], torch.zeros_like(z[3]))
                l_reg_dec = torch.zeros(1, device=x6.device)
                if optimize_decoder:
                    for k, v in cur.items():
                        l_reg_dec = l_reg_dec + l2_criterion(v, orig[k])
                total = hp["rec_6d_w"] * l6 + hp["rec_rot_w"] * lrot + hp["rec_pose_w"] * lpos + hp["reg_w"] * l_reg \
                    + hp["reg_w_decoder"] * l_reg_dec
                hist.append(torch.stack([l6.detach(), lrot.detach(), lpos.detach(), l_reg.detach(), l_reg_dec.detach().reshape(()),
                                         total.detach().reshape(())]))
                dec_phase = optimize_decoder and i > prev_epochs
                (d_opt if dec_phase else z_opt).zero_grad()
                total.reshape(()).backward()
                if dec_phase:
                    d_opt.step()
                    if d_sched is not None:
                        d_sched.step()
                else:
                    z_opt.step()
                    if z_sched is not None:
                        z_sched.step()
        finally:
            self.params = saved
        return dict(losses=torch.stack(hist), out_6d=x6.detach(), out_rot_mat=rot.detach(), out_pose_pos=pos.detach(),
                    z=[v.detach() for v in z], decoder_params={k: v.detach() for k, v in cur.items()})

    def test_path(self, seq_rot_6d, seq_rot_mat, sampled_z):
        """TwoHierSAVAEModel.test, seq_two_hier_sa_vae.py:560-639 (random_root_rot_flag False).

        Returns (gt_pos, mean_pos, sampled_pos) as [T, B, 24, 3]."""
        bs, t, _ = seq_rot_6d.shape
        j = len(self.parents)
        with torch.no_grad():
            gt = forward_kinematics(seq_rot_mat.view(bs * t, j, 3, 3), self.parents, self.offsets).view(bs, t, j, 3)
            _, zs = self.encode(seq_rot_6d.view(bs, t, -1).transpose(1, 2))
            mean_z = []
            for zi, dist in enumerate(zs):
                d = self.hp["shallow_latent_d"] if zi == 0 else self.hp["latent_d"]
                mean_z.append(dist[:, :, :d])
            _, _, mpos = self.decode(mean_z)
            _, _, spos = self.decode(sampled_z)
        return gt.transpose(0, 1), mpos.view(bs, t, j, 3).transpose(0, 1), spos.view(bs, t, j, 3).transpose(0, 1)


class TrajectoryOracle:
    """Restatement of trajectory_pred_model.py Encoder (:45-115) + TrajectoryModel.forward (:206-260)."""

    def __init__(self, hp, mean_std, parents=topo.SMPL24_PARENTS):
        self.hp = hp
        self.n = hp["num_layers"]
        self.levels = topo.hierarchy(parents, self.n, hp["skeleton_dist"])
        self.edge_num = [len(l["edges"]) for l in self.levels]
        base = 3 if hp["trajectory_input_joint_pos"] else 6
        self.cbase = [base * 2 ** i for i in range(self.n + 1)]
        self.ksize = hp["kernel_size"]
        self.pad = (self.ksize - 1) // 2
        ms = mean_std.clone()
        ms[1, ms[1] == 0] = 1.0
        self.mean = ms[0].float()
        self.std = ms[1].float()
        self.params, self.masks = {}, {}

    def init(self, seed=0):
        torch.manual_seed(seed)
        p = {}
        for i in range(self.n):
            cin, cout = self.cbase[i] * self.edge_num[i], self.cbase[i + 1] * self.edge_num[i]
            w, b = _init_skeleton_conv(self.levels[i]["neighbours"], cin, cout, self.ksize, True)
            p["enc.layers.%d.0.weight" % i], p["enc.layers.%d.0.bias" % i] = w, b
        p["fc_mapping.weight"], p["fc_mapping.bias"] = _init_linear(self.cbase[-1] * 7, 3)
        _kaiming_linear_(p["fc_mapping.weight"], p["fc_mapping.bias"])
        self.set_params(p)
        return self

    def set_params(self, p, device=None):
        self.params = {k: v.detach().clone().to(device or v.device).requires_grad_(True) for k, v in p.items()
                       if not k.endswith("mask")}
        dev = next(iter(self.params.values())).device
        for i in range(self.n):
            cin, cout = self.cbase[i] * self.edge_num[i], self.cbase[i + 1] * self.edge_num[i]
            self.masks["enc.layers.%d.0" % i] = conv_mask(self.levels[i]["neighbours"], cin, cout, self.ksize).to(dev)
        self.mean, self.std = self.mean.to(dev), self.std.to(dev)

    def accumulate(self, pose, root_v):
        """gen_motion_w_trajectory, trajectory_pred_model.py:289-303.  pose [T,B,24,3], root_v [T,B,3] (standardised)."""
        v = self.mean[576:579] + self.std[576:579] * root_v
        v = torch.cat([torch.zeros_like(v[:1]), v[1:]], dim=0)
        return pose + torch.cumsum(v, dim=0)[:, :, None, :]

    def step(self, seq_rot_pos, seq_joint_pos, seq_root_v, backward=True):
        hp, p = self.hp, self.params
        bs, t, _ = seq_joint_pos.shape
        x = seq_joint_pos.transpose(1, 2)
        for i in range(self.n):
            key = "enc.layers.%d.0" % i
            x = skeleton_conv(x, p[key + ".weight"], self.masks[key], p[key + ".bias"], 1, self.pad, hp["padding_mode"])
            x = skeleton_pool(x, self.levels[i]["pooling_list"], self.cbase[i + 1])
            x = F.leaky_relu(x, 0.2)
        d = self.cbase[-1]
        k_edges = x.shape[1] // d
        feat = x.view(bs, k_edges, d, t).transpose(2, 3).transpose(1, 2).reshape(bs, t, -1)
        root_v_out = F.linear(feat, p["fc_mapping.weight"], p["fc_mapping.bias"])
        l_v = l2_criterion(root_v_out, seq_root_v)
        if hp["use_accumulation_root_v"]:
            pose = seq_rot_pos.view(bs, t, 24, 3).transpose(0, 1)
            pred = self.accumulate(pose, root_v_out.transpose(0, 1))
            gt = self.accumulate(pose, seq_root_v.transpose(0, 1))
            l_t = l2_criterion(pred, gt)
        else:
            l_t = torch.zeros((), device=x.device)
        total = hp["rec_root_v_w"] * l_v + hp["rec_root_trans_w"] * l_t
        if backward:
            total.backward()
        return dict(total=total, rec_root_v=l_v, rec_root_trans=l_t, root_v_out=root_v_out)


# --------------------------------------------------------------------------------------
# synthetic inputs
# --------------------------------------------------------------------------------------


def synthetic_batch(bs, t, parents, offsets, seed=1234, device="cpu", mean_std=None):
    g = torch.Generator().manual_seed(seed)
    x6 = torch.randn(bs, t, len(parents), 6, generator=g)
    rot = rot6d_to_rotmat(x6)
    seq_rot_mat = rot.reshape(bs, t, -1)
    seq_rot_6d = rotmat_to_rot6d(rot).reshape(bs, t, -1)
    pos = forward_kinematics(rot.view(bs * t, len(parents), 3, 3), parents, offsets).view(bs, t, -1)
    out = dict(seq_rot_6d=seq_rot_6d, seq_rot_mat=seq_rot_mat, seq_rot_pos=pos,
               seq_root_v=torch.randn(bs, t, 3, generator=g))
    if mean_std is not None:
        ms = mean_std.clone().float()
        ms[1, ms[1] == 0] = 1.0
        out["seq_joint_pos"] = (pos - ms[0, 360:432]) / ms[1, 360:432]
    return {k: v.to(device) for k, v in out.items()}


def draw_eps(oracle, bs, seed=4321, device="cpu"):
    g = torch.Generator().manual_seed(seed)
    hp = oracle.hp
    out = []
    for zi in range(oracle.n):
        d = hp["shallow_latent_d"] if zi == 0 else hp["latent_d"]
        k_edges = len(oracle.levels[zi]["pooling_list"])
        out.append(torch.randn(bs * k_edges, d, generator=g).to(device))
    return out
