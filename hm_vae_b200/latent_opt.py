"""Latent-space optimisation loop of the reference's downstream tasks, on the B200 kernels.

The reference embeds the same inner loop in ``final_long_seq_try_interpolation`` (seq_two_hier_sa_vae.py:1356-1429) and
``final_motion_completion_long_seq`` (:1698-1757), between file IO and visualisation (both out of scope):

    for i in range(opt_it):                                      # configs/len_64_test_interpolation.yaml:68-84 -> 150
        out = _decode_w_given_decoder(z_vec_list, curr_decoder)  # decoder -> 6D -> R -> FK      (:501-529)
        l_rec_* = l2_masked_criterion(out_*, target_*, mask)     # temporal / joint mask          (:717-735)
        l_reg = l2(z[0], 0) + l2(z[3], 0);  l_reg_decoder = sum_p l2(p, dec.state_dict()[name])
        l_total = w6d l6 + wrot lrot + wpos lpos + reg_w l_reg + reg_w_decoder l_reg_decoder
        i <= prev_epochs: Adam(z_vec_list, lr=opt_lr).step()     else: Adam(curr_decoder.parameters(), lr=opt_lr*1e-3).step()
        (+ StepLR(opt_step_size, opt_gamma) of whichever optimiser stepped)

Here one iteration is: decoder forward (tcgen05 conv kernels) -> ``hmvae_recon_masked_fwdbwd`` (GT FK, rot6d->R, FK, the three
masked MSE sums and d(total)/d(decoder output) in ONE kernel) -> decoder backward (dgrad only while the latents are optimised,
dgrad + wgrad while the decoder copy is) -> ``hmvae_l2_reg_fwdbwd`` (all regulariser terms and their gradients, one launch) ->
fused multi-tensor Adam with the StepLR position on the device.  Each of the two phases is captured once as a CUDA graph and
replayed; loss values stay on the device ([opt_it, 6] history tensor), nothing synchronises with the host.
"""
import copy

import torch

from . import ops


def l2_masked_criterion(pred, gt, mask):
    """seq_two_hier_sa_vae.py:717-735: mean over ALL elements of (pred-gt)^2 * mask; also the per-frame mean [bs, T].
    pred / gt: bs X T X 24 X {6 | 3 X 3 | 3}; mask: bs X T X 24."""
    assert pred.size() == gt.size()
    m = mask
    while m.dim() < pred.dim():
        m = m[..., None]
    loss = (pred - gt) ** 2 * m
    bs, timesteps = loss.shape[0], loss.shape[1]
    return loss.mean(), loss.reshape(bs, timesteps, -1).mean(dim=-1)


def decode_w_given_decoder(model, z_list, curr_decoder):
    """seq_two_hier_sa_vae.py:501-529 (``_decode`` with an explicit decoder copy)."""
    result = curr_decoder(z_list)
    bs = result.size(0)
    decoder_out = ops.transpose_ct(result).view(bs * model.max_timesteps, model.n_joints, -1)
    cont6d_rep = decoder_out[:, :, :model.output_dim]
    rot = ops.rot6d_to_rotmat(cont6d_rep)
    pos = model.fk_layer(rot)
    return (cont6d_rep.view(bs, model.max_timesteps, model.n_joints, -1),
            rot.view(bs, model.max_timesteps, model.n_joints, 3, 3),
            pos.view(bs, model.max_timesteps, model.n_joints, 3), None, None, None, None)


class LatentOptimizer:
    """State of one optimisation problem (one window): latents, decoder copy, the two optimisers, captured graphs."""

    LOSS_NAMES = ("rec_6d", "rec_rot", "rec_pose", "reg", "reg_decoder", "total")

    def __init__(self, model, target_cont6d, target_rotmat, target_mask, hp, z_vec_list=None, prev_epochs=50):
        self.model, self.hp, self.prev_epochs = model, hp, int(prev_epochs)
        dev = model.mean_vals.device
        f32 = dict(device=dev, dtype=torch.float32)
        bs, t = target_cont6d.shape[0], target_cont6d.shape[1]
        j = model.n_joints
        assert t == model.max_timesteps
        self.bs, self.t = bs, t
        self.gt6 = target_cont6d.to(**f32).reshape(bs, t, j * 6).contiguous()
        self.gtR = target_rotmat.to(**f32).reshape(bs, t, j * 9).contiguous()
        self.mask = target_mask.to(**f32).reshape(bs, t, j).contiguous()
        n = hp['num_layers']
        k_edges = [len(p) for p in model.enc.pooling_list]
        lat = [model.shallow_latent_d] + [model.latent_d] * (n - 1)
        if z_vec_list is None:      # :1318-1331: N(0,1) shallow / deep latents, zeros for the two dead levels
            z_vec_list = [torch.randn(bs, k_edges[i], lat[i], **f32) if i in (0, n - 1) else torch.zeros(bs, k_edges[i], lat[i], **f32)
                          for i in range(n)]
        self.z_vec_list = [torch.nn.Parameter(z.detach().to(**f32).clone()) for z in z_vec_list]
        self.live_z = [self.z_vec_list[0], self.z_vec_list[n - 1]] if n > 1 else [self.z_vec_list[0]]
        self.z_zero = [torch.zeros_like(z) for z in self.live_z]
        self.optimize_decoder = bool(hp.get('optimize_decoder', False))
        wd = hp['weight_decay']
        self.z_opt = ops.FusedAdam(self.z_vec_list, lr=hp['opt_lr'], weight_decay=wd)
        if hp.get('opt_lr_policy', 'constant') == 'step':
            self.z_opt.set_schedule(hp['opt_gamma'], hp['opt_step_size'])
        self.decoder = model.dec
        if self.optimize_decoder:
            self.decoder = copy.deepcopy(model.dec)                        # includes the dec.enc.* copies, like the reference
            self.dec_named = [(k, p) for k, p in self.decoder.named_parameters() if p.requires_grad]
            ref = dict(model.dec.named_parameters())
            self.dec_params = [p for _, p in self.dec_named]
            self.dec_refs = [ref[k].detach() for k, _ in self.dec_named]
            self.dec_opt = ops.FusedAdam(self.dec_params, lr=hp['opt_lr'] * 0.001, weight_decay=wd)
            if hp.get('opt_lr_policy', 'constant') == 'step':
                self.dec_opt.set_schedule(hp['opt_gamma'], hp['opt_step_size'])
        self.acc = torch.zeros(8, **f32)
        self.res = torch.zeros(8, **f32)
        self.out6 = torch.empty(bs, model.n_joints * 6, t, **f32)          # decoder output of the last iteration (NCW)
        self.rot = torch.empty(bs, t, j, 3, 3, **f32)
        self.pos = torch.empty(bs, t, j, 3, **f32)
        self._graphs = {}
        self.it = 0

    # ------------------------------------------------------------------ one iteration (capturable)
    def _iteration(self, phase):
        hp, model = self.hp, self.model
        dec_phase = phase == "decoder"
        for z in self.live_z:
            z.requires_grad_(not dec_phase)
            z.grad = None
        if self.optimize_decoder:
            for p in self.dec_params:
                p.requires_grad_(dec_phase)
                p.grad = None
        n = hp['num_layers']
        z_in = [None] * n
        z_in[0], z_in[n - 1] = self.z_vec_list[0], self.z_vec_list[n - 1]
        if dec_phase:
            ops.prefetch_packs(self.decoder.conv_plans())
        out = self.decoder(z_in)                                           # bs X (24*6) X T
        self.out6.copy_(out.detach())
        fk_off = model.fk_layer.positions[0].contiguous()
        dx6 = ops.recon_fwdbwd(out.detach(), True, self.gt6, self.gtR, fk_off, model._parents, hp['rec_6d_w'], hp['rec_rot_w'],
                               hp['rec_pose_w'], self.acc, want_grad=True, mask=self.mask, rot_out=self.rot, pos_out=self.pos)
        with ops.wgrad_overlap():
            out.backward(dx6)
        # regularisers: value every iteration (the reference prints them), gradient for whichever side is being optimised
        reg_w = float(hp.get('reg_w', 0.0))
        zg = None
        if not dec_phase and reg_w != 0.0:
            zg = [z.grad for z in self.live_z]
        ops.l2_reg_fwdbwd(self.live_z, self.z_zero, reg_w, self.acc[3:4], grads=zg, accumulate=[True] * len(self.live_z))
        if self.optimize_decoder:
            grads = acc_flags = None
            if dec_phase:
                grads, acc_flags = [], []
                for p in self.dec_params:
                    acc_flags.append(p.grad is not None)
                    if p.grad is None:
                        p.grad = torch.empty_like(p)                       # reached by the regulariser only (the dec.enc.* copies)
                    grads.append(p.grad)
            ops.l2_reg_fwdbwd(self.dec_params, self.dec_refs, float(hp.get('reg_w_decoder', 0.0)), self.acc[4:5], grads=grads,
                              accumulate=acc_flags)
        nf, j = float(self.bs * self.t), model.n_joints
        scale = [1.0 / (nf * 6 * j), 1.0 / (nf * 9 * j), 1.0 / (nf * 3 * j), 1.0, 1.0]
        w_all = [hp['rec_6d_w'], hp['rec_rot_w'], hp['rec_pose_w'], reg_w, float(hp.get('reg_w_decoder', 0.0))]
        ops.loss_finalize(self.acc, self.res, scale, w_all, [0.0] * 5)      # res = [l6, lrot, lpos, l_reg, l_reg_dec, total, -]
        (self.dec_opt if dec_phase else self.z_opt).step_dyn()

    def _phase(self, i):
        return "decoder" if (self.optimize_decoder and i > self.prev_epochs) else "z"

    def run(self, opt_it=None, cuda_graph=True):
        """Runs ``opt_it`` iterations (default hp['opt_it']).  Returns the [opt_it, 6] device tensor of per-iteration losses
        (LOSS_NAMES order); outputs of the last iteration are in ``result()``."""
        opt_it = int(self.hp['opt_it'] if opt_it is None else opt_it)
        hist = torch.zeros(opt_it, 6, device=self.acc.device, dtype=torch.float32)
        for k in range(opt_it):
            i = self.it
            phase = self._phase(i)
            (self.dec_opt if phase == "decoder" else self.z_opt).advance()
            if cuda_graph:
                g = self._graphs.get(phase)
                if g is None:
                    g = self._capture(phase)
                    if g is None:                      # the eager warm-up inside _capture WAS this iteration
                        hist[k].copy_(self.res[:6])
                        self.it += 1
                        continue
                g.replay()
            else:
                self._iteration(phase)
            hist[k].copy_(self.res[:6])
            self.it += 1
        return hist

    def _capture(self, phase):
        """First iteration of a phase runs eagerly (allocates the plans / packed weights / workspaces), the second one is captured.
        Returns None after the eager iteration, the graph once captured."""
        key = ("warm", phase)
        if key not in self._graphs:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                self._iteration(phase)
            torch.cuda.current_stream().wait_stream(side)
            self._graphs[key] = True
            return None
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        ops._force_repack = phase == "decoder"       # the packed tf32 weight copies change every replay in the decoder phase
        try:
            with torch.cuda.graph(graph):
                self._iteration(phase)
        finally:
            ops._force_repack = False
        self._graphs[phase] = graph
        graph.replay()                                # capture executes nothing: this replay IS the iteration
        return _Done()

    def result(self):
        bs, t, j = self.bs, self.t, self.model.n_joints
        out6 = ops.transpose_ct(self.out6).view(bs, t, j, 6)
        return dict(out_6d=out6, out_rot_mat=self.rot, out_pose_pos=self.pos, z_vec_list=[z.detach() for z in self.z_vec_list],
                    decoder=self.decoder)


class _Done:
    """Returned by ``_capture`` right after it replayed the freshly captured graph once."""

    def replay(self):
        pass


def optimize_latent(model, target_cont6d, target_rotmat, target_mask, hp, z_vec_list=None, prev_epochs=50, opt_it=None,
                    cuda_graph=True):
    """The inner loop of seq_two_hier_sa_vae.py:1356-1429 / :1698-1757 for one window.  Returns (result dict, loss history)."""
    lo = LatentOptimizer(model, target_cont6d, target_rotmat, target_mask, hp, z_vec_list=z_vec_list, prev_epochs=prev_epochs)
    hist = lo.run(opt_it, cuda_graph=cuda_graph)
    res = lo.result()
    res["losses"] = hist
    return res
