"""Generates tests/golden/* and hm_vae_b200/data/smpl24.npz by RUNNING THE REAL REFERENCE.

Run in the authoring container only (needs /root/reference; the GPU box never runs this):

    python oracle/make_golden.py

What runs the reference's own code:
  * skeleton.py (get_edges / find_neighbor / SkeletonPool / SkeletonConv / SkeletonUnpool), fk_layer.py,
    my_tools.py -- imported unmodified.
  * seq_two_hier_sa_vae.py / trajectory_pred_model.py / trainer_motion_vae.weights_init -- imported under
    sys.modules stubs for the packages that are absent (torchgeometry, utils_common's matplotlib deps,
    lib.utils.eval_utils).  ``TwoHierSAVAEModel.forward`` and ``TrajectoryModel.forward`` are executed AS
    WRITTEN on CPU by neutralising ``Tensor.cuda`` and injecting epsilon through ``torch.randn_like``
    (the objects are built without running their ``__init__``, which hard-codes absolute paths).

Nothing here is copied into the product; only the numeric outputs are stored.
"""
import json
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn
import yaml

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def import_reference():
    sys.path.insert(0, REF)
    noop = lambda *a, **k: None
    _stub("torchgeometry")
    _stub("utils_common", show3Dpose_animation=noop, show3Dpose_animation_multiple=noop,
          show3Dpose_animation_with_mask=noop, write_loss=noop, write_images=noop, write_images_interpolation=noop)
    _stub("lib")
    _stub("lib.utils")
    _stub("lib.utils.eval_utils", compute_accel=noop, compute_error_accel=noop, compute_error_verts=noop,
          batch_compute_similarity_transform_torch=noop)
    try:
        import torch.utils.tensorboard  # noqa: F401
    except Exception:
        _stub("torch.utils.tensorboard", SummaryWriter=object)
    import skeleton, fk_layer, my_tools, seq_two_hier_sa_vae, trajectory_pred_model, trainer_motion_vae
    return skeleton, fk_layer, my_tools, seq_two_hier_sa_vae, trajectory_pred_model, trainer_motion_vae


def to_np(d):
    return {k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in d.items()}


def checksum(t):
    t = t.detach().double()
    return [float(t.sum()), float(t.abs().sum()), float((t * t).sum())]


def main():
    os.makedirs(GOLD, exist_ok=True)
    skeleton, fk_layer, my_tools, hm, traj, trainer = import_reference()
    torch.set_num_threads(8)

    # ------------------------------------------------------------------ data fixtures
    parents = json.load(open(f"{REF}/utils/data/joint24_parents.json"))
    offsets = np.load(f"{REF}/utils/data/skeleton_offsets.npy")
    rest = np.load(f"{REF}/utils/data/rest_pose_coord.npy")
    mean_std = np.load(f"{REF}/utils/data/for_all_data_motion_model/all_amass_data_mean_std.npy")
    os.makedirs(f"{ROOT}/hm_vae_b200/data", exist_ok=True)
    np.savez(f"{ROOT}/hm_vae_b200/data/smpl24.npz", parents=np.asarray(parents, np.int32), offsets=offsets,
             rest_pose=rest, mean_std=mean_std)

    # ------------------------------------------------------------------ topology (bit-exact ints)
    edges = skeleton.get_edges(f"{REF}/utils/data/joint24_parents.json")
    levels = []
    for i in range(4):
        nb = skeleton.find_neighbor(edges, 2)
        pool = skeleton.SkeletonPool(edges, "mean", 2, last_pool=(i == 3))
        levels.append(dict(edges=[list(e) for e in edges], neighbours=nb, seq_list=pool.seq_list,
                           pooling_list=pool.pooling_list, new_edges=[list(e) for e in pool.new_edges],
                           edge_mat=skeleton.calc_edge_mat(edges)))
        edges = pool.new_edges
    json.dump(dict(parents=parents, levels=levels), open(f"{GOLD}/topology.json", "w"))

    # ------------------------------------------------------------------ module-level vectors
    g = torch.Generator().manual_seed(7)
    mod = {}
    fk = fk_layer.ForwardKinematicsLayer(device=torch.device("cpu"), parents=parents, positions=offsets)
    ident = torch.eye(3)[None, None].repeat(1, 24, 1, 1)
    mod["fk_identity"] = fk(ident)
    x6 = torch.randn(5, 24, 6, generator=g)
    x6[1, 3] = 0.0            # zero input -> zero matrix (no NaN), my_tools.py:8 eps clamp
    x6[2, 5, 3:] = x6[2, 5, :3] * 2.0   # parallel vectors -> degenerate cross product
    x6r = x6.clone().requires_grad_(True)
    rot = my_tools.rotation_matrix_from_ortho6d(x6r)
    gr = torch.randn(5, 24, 3, 3, generator=g)
    rot.backward(gr)
    mod.update(rot6d_x=x6, rot6d_R=rot, rot6d_gR=gr, rot6d_gx=x6r.grad)
    rin = rot.detach().clone().requires_grad_(True)
    pos = fk(rin)
    gp = torch.randn(5, 24, 3, generator=g)
    pos.backward(gp)
    mod.update(fk_R=rin, fk_pos=pos, fk_gpos=gp, fk_gR=rin.grad)
    # non-orthonormal "rotations": FK must not assume orthonormality
    rr = torch.randn(4, 24, 3, 3, generator=g).requires_grad_(True)
    pos2 = fk(rr)
    gp2 = torch.randn(4, 24, 3, generator=g)
    pos2.backward(gp2)
    mod.update(fk2_R=rr, fk2_pos=pos2, fk2_gpos=gp2, fk2_gR=rr.grad)
    x6b = torch.randn(3, 24, 6, generator=g).requires_grad_(True)
    pos3 = fk(x6b)
    gp3 = torch.randn(3, 24, 3, generator=g)
    pos3.backward(gp3)
    mod.update(fk6_x=x6b, fk6_pos=pos3, fk6_gpos=gp3, fk6_gx=x6b.grad)
    # custom per-frame positions argument (fk_layer.py:82-89)
    cpos = torch.randn(4, 24, 3, generator=g)
    mod.update(fkp_positions=cpos, fkp_pos=fk(rr.detach(), cpos))

    # SkeletonConv cases: (level, ci, co, K, stride, pad, mode, bias, B, T)
    cases = [(0, 2, 4, 3, 1, 1, "reflection", True, 2, 8),
             (0, 2, 4, 15, 2, 7, "reflection", True, 2, 16),
             (1, 4, 2, 5, 1, 2, "zeros", False, 3, 9),
             (2, 4, 8, 15, 2, 7, "reflection", True, 2, 8),
             (3, 8, 4, 15, 1, 7, "reflection", True, 2, 8)]
    for n, (lvl, ci, co, k, s, p, mode, bias, b, t) in enumerate(cases):
        torch.manual_seed(100 + n)
        nb = levels[lvl]["neighbours"]
        j = len(nb)
        conv = skeleton.SkeletonConv(nb, j * ci, j * co, k, j, stride=s, padding=p, bias=bias, padding_mode=mode)
        x = torch.randn(b, j * ci, t).requires_grad_(True)
        y = conv(x)
        gy = torch.randn_like(y)
        y.backward(gy)
        mod.update({f"conv{n}_cfg": np.asarray([lvl, ci, co, k, s, p, int(mode == "reflection"), int(bias), b, t]),
                    f"conv{n}_w": conv.weight, f"conv{n}_mask_sum": conv.mask.sum(),
                    f"conv{n}_x": x, f"conv{n}_y": y, f"conv{n}_gy": gy, f"conv{n}_gx": x.grad,
                    f"conv{n}_gw": conv.weight.grad})
        if bias:
            mod.update({f"conv{n}_b": conv.bias, f"conv{n}_gb": conv.bias.grad})
    # pool / unpool / upsample
    for lvl in range(4):
        c = 3
        pl = skeleton.SkeletonPool([tuple(e) for e in levels[lvl]["edges"]], "mean", c, last_pool=(lvl == 3))
        x = torch.randn(2, len(levels[lvl]["edges"]) * c, 5, generator=g)
        y = pl(x)
        un = skeleton.SkeletonUnpool(pl.pooling_list, c)
        mod.update({f"pool{lvl}_x": x, f"pool{lvl}_y": y, f"unpool{lvl}_y": un(y),
                    f"pool{lvl}_w": pl.weight, f"unpool{lvl}_w": un.weight})
    xu = torch.randn(2, 6, 7, generator=g)
    mod.update(up_x=xu, up_y=nn.Upsample(scale_factor=2, mode="linear", align_corners=False)(xu))
    np.savez_compressed(f"{GOLD}/modules.npz", **to_np(mod))

    # ------------------------------------------------------------------ model-level vectors
    from oracle import hmvae_ref as O

    torch.Tensor.cuda = lambda self, *a, **k: self      # run the reference forward on CPU, unmodified
    off_t = torch.from_numpy(offsets)
    out = {}
    for tag, cfg_name, bs in [("len64", "len64_no_aug_hm_vae.yaml", 2), ("len8", "len8_data_aug_hm_vae.yaml", 3)]:
        hp = yaml.safe_load(open(f"{REF}/configs/{cfg_name}"))
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
        holder.apply(trainer.weights_init(hp["init"]))       # what Trainer.__init__ does (trainer_motion_vae.py:35)

        ora = O.HMVAEOracle(hp, parents, off_t)
        batch = O.synthetic_batch(bs, hp["train_seq_len"], parents, off_t, seed=1234)
        eps = O.draw_eps(ora, bs, seed=4321)
        data = (batch["seq_rot_6d"], batch["seq_rot_mat"], batch["seq_rot_pos"], batch["seq_rot_pos"],
                batch["seq_rot_pos"], batch["seq_rot_pos"], batch["seq_root_v"])
        sd = model.state_dict()
        out[f"{tag}_nkeys"] = len(sd)
        out[f"{tag}_keys"] = np.asarray(sorted(sd.keys()))
        for k, v in sd.items():
            if k.startswith("dec.enc."):
                continue
            out[f"{tag}_init/{k}"] = np.asarray(checksum(v))
        for it_tag, iters in [("it0", 0), ("itlate", hp["iteration_interval"] + 1)]:
            model.zero_grad()
            q = list(eps)
            orig = torch.randn_like
            torch.randn_like = lambda t, *a, **k: q.pop(0)
            try:
                res = model(data, hp, iters)
            finally:
                torch.randn_like = orig
            names = ["total", "kl", "rec_6d", "rec_rot", "rec_pose"]
            out[f"{tag}_{it_tag}_losses"] = np.asarray([float(res[i]) for i in range(5)] +
                                                       [float(res[9][0]), float(res[9][3])])
            for k, p in model.named_parameters():
                if k.startswith("dec.enc.") or not p.requires_grad:
                    continue
                out[f"{tag}_{it_tag}_grad/{k}"] = np.asarray(checksum(p.grad) if p.grad is not None else [np.nan] * 3)
            if it_tag == "it0":
                out[f"{tag}_gb_enc0"] = model.enc.layers[0][0].bias.grad.clone().numpy()
                out[f"{tag}_gb_dec3"] = model.dec.convs[3].bias.grad.clone().numpy()
        # test() path (inference, config 4)
        hp2 = dict(hp)
        hp2["random_root_rot_flag"] = False
        sz = [torch.randn(bs, len(ora.levels[i]["pooling_list"]),
                          hp["shallow_latent_d"] if i == 0 else hp["latent_d"], generator=g) for i in range(4)]
        q = list(sz)
        orig = torch.randn_like
        torch.randn_like = lambda t, *a, **k: q.pop(0)
        try:
            gt_pos, mean_pos, samp_pos, _ = model.test(data, hp2, 0)
        finally:
            torch.randn_like = orig
        for i in range(4):
            out[f"{tag}_test_z{i}"] = sz[i].numpy()
        out[f"{tag}_test_gt"] = gt_pos.numpy()
        out[f"{tag}_test_mean"] = mean_pos.detach().numpy()
        out[f"{tag}_test_sampled"] = samp_pos.detach().numpy()

    # trajectory model
    hp = yaml.safe_load(open(f"{REF}/configs/trajectory_model.yaml"))
    torch.manual_seed(0)
    tm = traj.TrajectoryModel.__new__(traj.TrajectoryModel)
    nn.Module.__init__(tm)
    tm.latent_d, tm.n_joints, tm.input_dim, tm.output_dim = hp["latent_d"], hp["n_joints"], hp["input_dim"], hp["output_dim"]
    tm.max_timesteps = hp["train_seq_len"]
    tm.fk_layer = fk_layer.ForwardKinematicsLayer(device=torch.device("cpu"), parents=parents, positions=offsets)
    tm.hp = hp
    tm.enc = traj.Encoder(hp, skeleton.get_edges(f"{REF}/utils/data/joint24_parents.json"))
    tm.d_model = tm.enc.channel_base[-1]
    tm.fc_mapping = nn.Linear(tm.d_model * 7, 3)
    ms = mean_std.copy()
    ms[1, ms[1, :] == 0] = 1.0
    tm.mean_vals = torch.from_numpy(ms[0, :]).float()[None, :]
    tm.std_vals = torch.from_numpy(ms[1, :]).float()[None, :]
    holder = nn.Module()
    holder.model = tm
    holder.apply(trainer.weights_init(hp["init"]))
    bs = 2
    batch = O.synthetic_batch(bs, hp["train_seq_len"], parents, off_t, seed=1234, mean_std=torch.from_numpy(mean_std))
    data = (batch["seq_rot_6d"], batch["seq_rot_mat"], batch["seq_rot_pos"], batch["seq_joint_pos"],
            batch["seq_rot_pos"], batch["seq_rot_pos"], batch["seq_root_v"])
    res = tm(data, hp, 0)
    out["traj_losses"] = np.asarray([float(res[0]), float(res[6]), float(res[8])])
    for k, v in tm.state_dict().items():
        out[f"traj_init/{k}"] = np.asarray(checksum(v))
    for k, p in tm.named_parameters():
        if p.requires_grad:
            out[f"traj_grad/{k}"] = np.asarray(checksum(p.grad) if p.grad is not None else [np.nan] * 3)
    out["traj_gb_fc"] = tm.fc_mapping.bias.grad.clone().numpy()
    np.savez_compressed(f"{GOLD}/models.npz", **to_np(out))
    print("golden written:", sorted(os.listdir(GOLD)))
    for f in os.listdir(GOLD):
        print(f, os.path.getsize(os.path.join(GOLD, f)))


if __name__ == "__main__":
    main()
