"""Pins the oracle (oracle/*.py) against vectors produced by the REAL reference (oracle/make_golden.py)."""
import numpy as np
import pytest
import torch

from oracle import hmvae_ref as O
from oracle import topology as topo

HP64 = dict(latent_d=24, shallow_latent_d=12, n_joints=24, input_dim=6, output_dim=6, num_layers=4, skeleton_dist=2,
            skeleton_pool="mean", extra_conv=0, padding_mode="reflection", kernel_size=15, upsampling="linear",
            train_seq_len=64, kl_w=0.003, shallow_kl_w=0.003, rec_6d_w=1, rec_rot_w=1, rec_pose_w=10,
            iteration_interval=50000)
HP8 = dict(HP64, latent_d=6, shallow_latent_d=6, kernel_size=3, train_seq_len=8, iteration_interval=20000)
HPT = dict(latent_d=12, n_joints=24, input_dim=6, output_dim=6, num_layers=4, skeleton_dist=2, skeleton_pool="mean",
           extra_conv=0, padding_mode="reflection", kernel_size=31, train_seq_len=128, trajectory_input_joint_pos=True,
           use_accumulation_root_v=True, rec_root_v_w=1, rec_root_trans_w=1)


def _norm(x):
    return [[list(e) if isinstance(e, (list, tuple)) else e for e in row] for row in x]


def test_topology_bit_exact(golden_topology):
    levels = topo.hierarchy(golden_topology["parents"], 4, 2)
    assert golden_topology["parents"] == topo.SMPL24_PARENTS
    for mine, ref in zip(levels, golden_topology["levels"]):
        assert [list(e) for e in mine["edges"]] == ref["edges"]
        assert mine["neighbours"] == ref["neighbours"]
        assert mine["seq_list"] == ref["seq_list"]
        assert mine["pooling_list"] == ref["pooling_list"]
        assert [list(e) for e in mine["new_edges"]] == ref["new_edges"]
        assert topo.edge_distance(mine["edges"]) == ref["edge_mat"]
    assert [len(topo.mask_blocks(l["neighbours"])) for l in levels] == [146, 92, 67, 49]


def test_pasted_cascade_docstring():
    """skeleton.py:464-477 (the author's pasted stdout) -- levels 1..3."""
    levels = topo.hierarchy()
    assert levels[0]["seq_list"] == [[0], [1, 4, 7, 10], [2, 5, 8, 11], [3, 6, 9], [12, 15], [13, 16, 18, 20, 22],
                                     [14, 17, 19, 21, 23]]
    assert levels[1]["pooling_list"] == [[0], [1, 2], [3, 4], [5, 6], [7], [8], [9, 10], [11], [12, 13]]
    assert _norm([levels[2]["new_edges"]])[0] == [[0, 24], [0, 10], [0, 11], [0, 9], [9, 15], [9, 22], [9, 23]]


def test_fk_known_answer(smpl, golden_modules):
    """FK(identity) == rest_pose_coord.npy and offsets[i] = rest[i] - rest[parent]."""
    off = torch.from_numpy(smpl["offsets"])
    eye = torch.eye(3)[None, None].repeat(1, 24, 1, 1)
    pos = O.forward_kinematics(eye, smpl["parents"], off)
    assert np.abs(pos[0].numpy() - smpl["rest_pose"]).max() < 1e-6
    np.testing.assert_allclose(pos.numpy(), golden_modules["fk_identity"], atol=1e-7)


def test_rot6d_and_fk_vs_reference(smpl, golden_modules):
    g = golden_modules
    off = torch.from_numpy(smpl["offsets"])
    x = torch.from_numpy(g["rot6d_x"]).requires_grad_(True)
    r = O.rot6d_to_rotmat(x)
    ok = np.ones((5, 24), bool)
    ok[2, 5] = False          # parallel a/b: z = normalize(rounding noise) -- ill-conditioned, only column x is defined
    np.testing.assert_allclose(r.detach().numpy()[ok], g["rot6d_R"][ok], atol=2e-6)
    np.testing.assert_allclose(r.detach().numpy()[2, 5, :, 0], g["rot6d_R"][2, 5, :, 0], atol=2e-6)
    assert not torch.isnan(r).any()
    assert float(r[1, 3].abs().max()) == 0.0       # zero 6D input -> zero matrix (eps clamp), no NaN
    r.backward(torch.from_numpy(g["rot6d_gR"]))
    np.testing.assert_allclose(x.grad.numpy()[ok], g["rot6d_gx"][ok], rtol=1e-4, atol=1e-5)
    for tag, key_in, key_g in [("fk", "fk_R", "fk_gR"), ("fk2", "fk2_R", "fk2_gR"), ("fk6", "fk6_x", "fk6_gx")]:
        rin = torch.from_numpy(g[key_in]).requires_grad_(True)
        pos = O.forward_kinematics(rin, smpl["parents"], off)
        np.testing.assert_allclose(pos.detach().numpy(), g[tag + "_pos"], rtol=1e-5, atol=1e-5)
        pos.backward(torch.from_numpy(g[tag + "_gpos"]))
        np.testing.assert_allclose(rin.grad.numpy(), g[key_g], rtol=1e-4, atol=2e-5)
    pos = O.forward_kinematics(torch.from_numpy(g["fk2_R"]), smpl["parents"], torch.from_numpy(g["fkp_positions"]))
    np.testing.assert_allclose(pos.numpy(), g["fkp_pos"], rtol=1e-5, atol=1e-5)


def test_conv_pool_unpool_upsample_vs_reference(golden_modules, golden_topology):
    g = golden_modules
    for n in range(5):
        lvl, ci, co, k, s, p, refl, bias, b, t = [int(v) for v in g[f"conv{n}_cfg"]]
        nb = golden_topology["levels"][lvl]["neighbours"]
        j = len(nb)
        mask = O.conv_mask(nb, j * ci, j * co, k)
        assert float(mask.sum()) == float(g[f"conv{n}_mask_sum"])
        w = torch.from_numpy(g[f"conv{n}_w"]).requires_grad_(True)
        x = torch.from_numpy(g[f"conv{n}_x"]).requires_grad_(True)
        bb = torch.from_numpy(g[f"conv{n}_b"]).requires_grad_(True) if bias else None
        y = O.skeleton_conv(x, w, mask, bb, s, p, "reflection" if refl else "zeros")
        np.testing.assert_allclose(y.detach().numpy(), g[f"conv{n}_y"], rtol=1e-5, atol=1e-5)
        y.backward(torch.from_numpy(g[f"conv{n}_gy"]))
        np.testing.assert_allclose(x.grad.numpy(), g[f"conv{n}_gx"], rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(w.grad.numpy(), g[f"conv{n}_gw"], rtol=1e-5, atol=1e-5)
    for lvl in range(4):
        pl = golden_topology["levels"][lvl]["pooling_list"]
        x = torch.from_numpy(g[f"pool{lvl}_x"])
        y = O.skeleton_pool(x, pl, 3)
        assert np.array_equal(y.numpy(), g[f"pool{lvl}_y"])
        assert np.array_equal(O.skeleton_unpool(y, pl, 3).numpy(), g[f"unpool{lvl}_y"])
        assert np.array_equal(O.pool_weight(pl, 3, len(golden_topology["levels"][lvl]["edges"])).numpy(), g[f"pool{lvl}_w"])
        assert np.array_equal(O.unpool_weight(pl, 3).numpy(), g[f"unpool{lvl}_w"])
    np.testing.assert_allclose(O.upsample2_linear(torch.from_numpy(g["up_x"])).numpy(), g["up_y"], atol=1e-6)


def _cks(t):
    t = t.detach().double()
    return np.asarray([float(t.sum()), float(t.abs().sum()), float((t * t).sum())])


@pytest.mark.parametrize("tag,hp,bs", [("len64", HP64, 2), ("len8", HP8, 3)])
def test_hmvae_step_vs_reference(tag, hp, bs, golden_models, smpl):
    g = golden_models
    off = torch.from_numpy(smpl["offsets"])
    ora = O.HMVAEOracle(hp, smpl["parents"].tolist(), off).init(seed=0)
    # seeded init reproduces the reference's RNG consumption (incl. the double re-draw of enc linears)
    for k, v in ora.params.items():
        np.testing.assert_allclose(_cks(v), g[f"{tag}_init/{k}"], rtol=1e-6, atol=1e-6, err_msg=k)
    batch = O.synthetic_batch(bs, hp["train_seq_len"], smpl["parents"].tolist(), off, seed=1234)
    eps = O.draw_eps(ora, bs, seed=4321)
    for it_tag, iters in [("it0", 0), ("itlate", hp["iteration_interval"] + 1)]:
        for p in ora.params.values():
            p.grad = None
        res = ora.step(batch["seq_rot_6d"], batch["seq_rot_mat"], eps, iterations=iters)
        ref = g[f"{tag}_{it_tag}_losses"]
        mine = [res["total"], hp["kl_w"] * res["kl_deep"] + hp["shallow_kl_w"] * res["kl_shallow"], res["rec_6d"],
                res["rec_rot"], res["rec_pose"], res["kl_shallow"], res["kl_deep"]]
        np.testing.assert_allclose([float(m) for m in mine], ref, rtol=2e-5)
        for k, p in ora.params.items():
            refc = g[f"{tag}_{it_tag}_grad/{k}"]
            if np.isnan(refc).all():
                assert p.grad is None or float(p.grad.abs().sum()) == 0.0, k
            else:
                np.testing.assert_allclose(_cks(p.grad), refc, rtol=2e-3, atol=1e-7, err_msg=k)
        if it_tag == "it0":
            np.testing.assert_allclose(ora.params["enc.layers.0.0.bias"].grad.numpy(), g[f"{tag}_gb_enc0"], rtol=1e-3, atol=1e-7)
    sz = [torch.from_numpy(g[f"{tag}_test_z{i}"]) for i in range(4)]
    gt, mean, samp = ora.test_path(batch["seq_rot_6d"], batch["seq_rot_mat"], sz)
    np.testing.assert_allclose(gt.numpy(), g[f"{tag}_test_gt"], atol=1e-5)
    np.testing.assert_allclose(mean.numpy(), g[f"{tag}_test_mean"], atol=2e-5)
    np.testing.assert_allclose(samp.numpy(), g[f"{tag}_test_sampled"], atol=2e-5)


def test_state_dict_key_contract(golden_models):
    """74 keys at len64: what a drop-in must expose."""
    assert int(golden_models["len64_nkeys"]) == 74
    keys = set(golden_models["len64_keys"].tolist())
    assert "enc.layers.0.0.mask" in keys and "dec.layers.3.2.weight" in keys and "dec.unpools.0.weight" in keys


def test_trajectory_step_vs_reference(golden_models, smpl):
    g = golden_models
    off = torch.from_numpy(smpl["offsets"])
    ms = torch.from_numpy(smpl["mean_std"])
    ora = O.TrajectoryOracle(HPT, ms, smpl["parents"].tolist()).init(seed=0)
    for k, v in ora.params.items():
        np.testing.assert_allclose(_cks(v), g[f"traj_init/{k}"], rtol=1e-6, atol=1e-6, err_msg=k)
    batch = O.synthetic_batch(2, 128, smpl["parents"].tolist(), off, seed=1234, mean_std=ms)
    res = ora.step(batch["seq_rot_pos"], batch["seq_joint_pos"], batch["seq_root_v"])
    np.testing.assert_allclose([float(res["total"]), float(res["rec_root_v"]), float(res["rec_root_trans"])],
                               g["traj_losses"], rtol=2e-5)
    for k, p in ora.params.items():
        np.testing.assert_allclose(_cks(p.grad), g[f"traj_grad/{k}"], rtol=2e-3, atol=1e-7, err_msg=k)
    np.testing.assert_allclose(ora.params["fc_mapping.bias"].grad.numpy(), g["traj_gb_fc"], rtol=1e-3, atol=1e-7)


def test_aa2rot_against_scipy():
    """torchgeometry is absent (parity unpinned): pin Rodrigues against scipy away from theta ~ 0."""
    from scipy.spatial.transform import Rotation

    g = torch.Generator().manual_seed(3)
    aa = torch.randn(64, 3, generator=g)
    r = O.angle_axis_to_rotation_matrix(aa)
    assert r.shape == (64, 4, 4)
    ref = Rotation.from_rotvec(aa.numpy().astype(np.float64)).as_matrix()
    np.testing.assert_allclose(r[:, :3, :3].numpy(), ref, atol=2e-5)
    small = O.angle_axis_to_rotation_matrix(torch.tensor([[1e-4, -2e-4, 3e-4]]))
    np.testing.assert_allclose(small[0, :3, :3].numpy(), [[1, -3e-4, -2e-4], [3e-4, 1, -1e-4], [2e-4, 1e-4, 1]], atol=1e-7)


# ------------------------------------------------------------------------------------------------ latent-space optimisation (8f-4)
HPOPT = dict(HP64, weight_decay=1e-4, opt_lr=0.1, opt_it=150, reg_w=0, reg_w_decoder=1000, opt_lr_policy="step", opt_step_size=50,
             opt_gamma=0.1, interpolation_window=5, optimize_decoder=True)


@pytest.fixture(scope="module")
def golden_latent_opt():
    import os

    return dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "latent_opt.npz")))


def latent_opt_problem(g, smpl):
    """Inputs of the golden run (oracle/make_golden_latent_opt.py): hp overrides, targets, mask, initial latents."""
    opt_it, step_size, prev_epochs = [int(v) for v in g["hp_overrides"]]
    hp = dict(HPOPT, opt_it=opt_it, opt_step_size=step_size)
    off = torch.from_numpy(smpl["offsets"])
    batch = O.synthetic_batch(2, 64, smpl["parents"].tolist(), off, seed=int(g["seed_batch"]))
    z_init = [torch.from_numpy(g[f"z_init{k}"]) for k in range(4)]
    return hp, prev_epochs, batch["seq_rot_6d"].view(2, 64, 24, 6), batch["seq_rot_mat"].view(2, 64, 24, 3, 3), \
        torch.from_numpy(g["target_mask"]), z_init


def test_l2_masked_criterion_vs_reference(golden_latent_opt):
    g = golden_latent_opt
    loss, saved = O.l2_masked_criterion(torch.from_numpy(g["lmc_pred"]), torch.from_numpy(g["lmc_gt"]), torch.from_numpy(g["lmc_mask"]))
    np.testing.assert_allclose(float(loss), float(g["lmc_loss"]), rtol=1e-6)
    np.testing.assert_allclose(saved.numpy(), g["lmc_saved"], rtol=1e-6, atol=1e-7)


def test_latent_optimisation_vs_reference(golden_latent_opt, smpl):
    """The oracle's restatement of the loop (seq_two_hier_sa_vae.py:1356-1429) against the loop body run with the REAL reference
    modules: losses of every iteration (latent phase, then decoder phase, StepLR boundaries inside both), final latents, final
    outputs, and the drift of every decoder-copy parameter."""
    g = golden_latent_opt
    hp, prev_epochs, t6, tR, mask, z_init = latent_opt_problem(g, smpl)
    ora = O.HMVAEOracle(hp, smpl["parents"].tolist(), torch.from_numpy(smpl["offsets"])).init(seed=0)
    res = ora.latent_optimise(z_init, t6, tR, mask, hp, prev_epochs=prev_epochs)
    np.testing.assert_allclose(res["losses"].numpy(), g["losses"], rtol=2e-4, atol=1e-9)
    for k in range(4):
        np.testing.assert_allclose(res["z"][k].numpy(), g[f"z_final{k}"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(res["out_6d"].numpy(), g["out_6d"], atol=5e-5)
    np.testing.assert_allclose(res["out_rot_mat"].numpy(), g["out_rot_mat"], atol=5e-5)
    np.testing.assert_allclose(res["out_pose_pos"].numpy(), g["out_pose_pos"], atol=5e-5)
    for k, v in res["decoder_params"].items():
        name = k[4:] if k.startswith("dec.") else "enc." + k[4:]        # curr_decoder's own naming: dec.X -> X, enc.X -> enc.X
        mine, ref = _cks(v), g[f"dec_final/{name}"]
        assert abs(mine[0] - ref[0]) <= 1e-6 * ref[1] + 1e-6, k          # the plain sum cancels: compare on the abs-sum scale
        np.testing.assert_allclose(mine[1:], ref[1:], rtol=1e-5, atol=1e-6, err_msg=k)
        np.testing.assert_allclose(_cks(v - ora.params[k].detach())[1:], g[f"dec_delta/{name}"][1:], rtol=5e-3, atol=1e-7, err_msg=k)
