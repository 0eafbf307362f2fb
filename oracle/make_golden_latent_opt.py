"""Generates tests/golden/latent_opt.npz by RUNNING THE REAL REFERENCE pieces of the latent-space optimisation loop
(seq_two_hier_sa_vae.py:1356-1429 / :1698-1757) on CPU.  Authoring container only (needs /root/reference):

    python oracle/make_golden_latent_opt.py

The loop lives inside two evaluation methods that read AMASS files from absolute paths and call matplotlib, so it cannot be
called as a whole.  What this script executes is the loop BODY with every operation done by the reference's own code:
``copy.deepcopy(model.dec)`` of the real Decoder, the real ``TwoHierSAVAEModel._decode_w_given_decoder`` (:501-529),
``l2_masked_criterion`` (:717-735), ``l2_criterion`` (:430-434), ``get_opt_scheduler`` (:40-51), torch.optim.Adam -- in the
order and with the zero_grad / step / scheduler.step rules of :1390-1423.  Only numeric results are stored.
"""
import copy
import json
import os
import sys

import numpy as np
import torch
import torch.nn as nn
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle.make_golden import GOLD, REF, checksum, import_reference, to_np  # noqa: E402


def build_reference_model(hm, skeleton, fk_layer, trainer, hp, parents, offsets):
    """The reference model without its __init__ (absolute paths), exactly as oracle/make_golden.py does."""
    torch.manual_seed(0)
    model = hm.TwoHierSAVAEModel.__new__(hm.TwoHierSAVAEModel)
    nn.Module.__init__(model)
    model.latent_d, model.shallow_latent_d = hp["latent_d"], hp["shallow_latent_d"]
    model.n_joints, model.input_dim, model.output_dim = hp["n_joints"], hp["input_dim"], hp["output_dim"]
    model.max_timesteps = hp["train_seq_len"]
    model.fk_layer = fk_layer.ForwardKinematicsLayer(device=torch.device("cpu"), parents=parents, positions=offsets)
    model.hp = hp
    model.enc = hm.Encoder(hp, skeleton.get_edges(f"{REF}/utils/data/joint24_parents.json"))
    model.dec = hm.Decoder(hp, model.enc)
    model.iteration_interval = hp["iteration_interval"]
    holder = nn.Module()
    holder.model = model
    holder.apply(trainer.weights_init(hp["init"]))
    return model


def main():
    skeleton, fk_layer, my_tools, hm, traj, trainer = import_reference()
    torch.set_num_threads(8)
    torch.Tensor.cuda = lambda self, *a, **k: self
    parents = json.load(open(f"{REF}/utils/data/joint24_parents.json"))
    offsets = np.load(f"{REF}/utils/data/skeleton_offsets.npy")
    hp = yaml.safe_load(open(f"{REF}/configs/len_64_test_interpolation.yaml"))
    hp = dict(hp, opt_it=7, opt_step_size=2)         # short run; StepLR boundaries inside both phases
    prev_epochs = 2                                  # iterations 0..2 optimise the latents, 3..6 the decoder copy
    model = build_reference_model(hm, skeleton, fk_layer, trainer, hp, parents, offsets)
    self = model

    from oracle import hmvae_ref as O
    bs, T = 2, hp["train_seq_len"]
    batch = O.synthetic_batch(bs, T, parents, torch.from_numpy(offsets), seed=777)
    input_cont6DRep = batch["seq_rot_6d"].view(bs, T, 24, 6)
    rotMatrices = batch["seq_rot_mat"].view(bs, T, 24, 3, 3)
    gt_fk_pose = model.fk_layer(rotMatrices.view(bs * T, 24, 3, 3)).view(bs, T, 24, 3).detach()
    temporal_mask = np.zeros(T)
    temporal_mask[::hp["interpolation_window"]] = 1           # :1299-1303
    temporal_mask[-1] = 1
    curr_target_mask = torch.from_numpy(temporal_mask).float()[None, :, None].repeat(bs, 1, 24)
    curr_target_mask[1, :, 13:] = 0.0                         # second sequence: a joint mask too (motion completion, :1640-1660)
    g = torch.Generator().manual_seed(99)
    z_init = [torch.randn(bs, 14, self.shallow_latent_d, generator=g), torch.zeros(bs, 9, self.latent_d),
              torch.zeros(bs, 7, self.latent_d), torch.randn(bs, 7, self.latent_d, generator=g)]
    self.z_vec_list = [nn.Parameter(z.clone()) for z in z_init]
    target_z_reg_list = [torch.zeros_like(z) for z in z_init]

    # D5: the loop calls ``curr_decoder(z_list, 1, 4)`` (:503) although Decoder.forward takes (z_vec_list, offset=None)
    # (:260) -- a stale signature.  The two extra positional arguments carried no tensor; drop them, run everything else as written.
    dec_forward = hm.Decoder.forward
    hm.Decoder.forward = lambda s, z, *stale: dec_forward(s, z)

    # ---- :1341-1349
    curr_decoder = copy.deepcopy(self.dec)
    self.gen_opt_for_decoder = torch.optim.Adam(list(curr_decoder.parameters()), lr=self.hp['opt_lr'] * 0.001,
                                                weight_decay=self.hp["weight_decay"])
    self.gen_scheduler_for_decoder = hm.get_opt_scheduler(self.gen_opt_for_decoder, self.hp)
    self.gen_opt = torch.optim.Adam(self.z_vec_list, lr=self.hp['opt_lr'], weight_decay=self.hp["weight_decay"])
    self.gen_scheduler = hm.get_opt_scheduler(self.gen_opt, self.hp)
    hist = []
    for i in range(self.hp["opt_it"]):                        # ---- :1356-1423
        opt_out_6d, opt_out_rot_mat, opt_out_pose_pos, _, _, _, _ = self._decode_w_given_decoder(self.z_vec_list, curr_decoder)
        l_rec_6d, _ = self.l2_masked_criterion(opt_out_6d, input_cont6DRep, curr_target_mask)
        l_rec_rot_mat, _ = self.l2_masked_criterion(opt_out_rot_mat, rotMatrices, curr_target_mask)
        l_rec_pose, _ = self.l2_masked_criterion(opt_out_pose_pos, gt_fk_pose, curr_target_mask)
        l_reg = self.l2_criterion(self.z_vec_list[0], target_z_reg_list[0]) + self.l2_criterion(self.z_vec_list[3], target_z_reg_list[3])
        l_reg_decoder = torch.zeros(1)
        for name, params in curr_decoder.named_parameters():
            l_reg_decoder += self.l2_criterion(params, self.dec.state_dict()[name])
        l_total = self.hp['rec_6d_w'] * l_rec_6d + self.hp['rec_rot_w'] * l_rec_rot_mat + self.hp['rec_pose_w'] * l_rec_pose + \
            self.hp['reg_w'] * l_reg + self.hp['reg_w_decoder'] * l_reg_decoder
        hist.append([float(l_rec_6d), float(l_rec_rot_mat), float(l_rec_pose), float(l_reg), float(l_reg_decoder), float(l_total)])
        if i > prev_epochs:
            self.gen_opt_for_decoder.zero_grad()
        else:
            self.gen_opt.zero_grad()
        l_total.backward()
        if i > prev_epochs:
            self.gen_opt_for_decoder.step()
            self.gen_scheduler_for_decoder.step()
        else:
            self.gen_opt.step()
            self.gen_scheduler.step()

    out = dict(hp_overrides=np.asarray([hp["opt_it"], hp["opt_step_size"], prev_epochs]), seed_batch=777, target_mask=curr_target_mask,
               losses=np.asarray(hist), out_6d=opt_out_6d, out_rot_mat=opt_out_rot_mat, out_pose_pos=opt_out_pose_pos)
    for k in range(4):
        out[f"z_init{k}"] = z_init[k]
        out[f"z_final{k}"] = self.z_vec_list[k]
    for name, p in curr_decoder.named_parameters():
        if p.requires_grad:
            out[f"dec_final/{name}"] = np.asarray(checksum(p))
            out[f"dec_delta/{name}"] = np.asarray(checksum(p - self.dec.state_dict()[name]))
    # l2_masked_criterion on its own (the three ranks it is called with)
    pm = torch.randn(2, 5, 24, 3, 3, generator=g)
    gm = torch.randn(2, 5, 24, 3, 3, generator=g)
    mk = (torch.rand(2, 5, 24, generator=g) > 0.5).float()
    l, sv = self.l2_masked_criterion(pm, gm, mk)
    out.update(lmc_pred=pm, lmc_gt=gm, lmc_mask=mk, lmc_loss=l, lmc_saved=sv)
    np.savez_compressed(f"{GOLD}/latent_opt.npz", **to_np(out))
    print("written", f"{GOLD}/latent_opt.npz", os.path.getsize(f"{GOLD}/latent_opt.npz"))
    print(np.asarray(hist))


if __name__ == "__main__":
    main()
