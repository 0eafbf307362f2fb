"""CPU-side checks of the host mirror: topology bookkeeping, state_dict contract, seeded init, C-ABI exports.
No kernel is launched here (no GPU in the authoring container)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import hm_vae_b200 as H
from hm_vae_b200 import _lib
from hm_vae_b200.seq_two_hier_sa_vae import TwoHierSAVAEModel
from hm_vae_b200.trajectory_pred_model import TrajectoryModel
from test_oracle_golden import HP64, HP8, HPT, _cks

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "hmvae_b200.h")).read()
    declared = set(re.findall(r"\b(hmvae_[a-z0-9_]+)\s*\(", header))
    declared.discard("hmvae_conv_plan")
    assert len(declared) >= 30
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), name
    assert set(_lib.EXPORTS) == declared
    assert lib.hmvae_version() >= 100


def test_topology_matches_golden(golden_topology):
    edges = H.get_edges(golden_topology["parents"])
    for i, ref in enumerate(golden_topology["levels"]):
        assert [list(e) for e in edges] == ref["edges"]
        assert H.calc_edge_mat(edges) == ref["edge_mat"]
        assert H.find_neighbor(edges, 2) == ref["neighbours"]
        pool = H.SkeletonPool(edges, "mean", 2, last_pool=(i == 3))
        assert pool.seq_list == ref["seq_list"] and pool.pooling_list == ref["pooling_list"]
        assert [list(e) for e in pool.new_edges] == ref["new_edges"]
        edges = pool.new_edges


def test_constructor_errors():
    nb = H.find_neighbor(H.get_edges([-1, 0, 0, 1]), 2)
    with pytest.raises(Exception, match="BAD"):
        H.SkeletonConv(nb, 9, 8, 3, 4)
    with pytest.raises(Exception, match="Unimplemented pooling mode"):
        H.SkeletonPool(H.get_edges([-1, 0, 0, 1]), "max", 2)
    with pytest.raises(ValueError):
        H.ForwardKinematicsLayer(device="cpu", parents=[-1, 2, 0], positions=np.zeros((3, 3), np.float32))


def test_mask_pool_unpool_buffers_match_golden(golden_modules, golden_topology):
    g = golden_modules
    for n in range(5):
        lvl, ci, co, k, s, p, refl, bias, b, t = [int(v) for v in g[f"conv{n}_cfg"]]
        nb = golden_topology["levels"][lvl]["neighbours"]
        conv = H.SkeletonConv(nb, len(nb) * ci, len(nb) * co, k, len(nb), stride=s, padding=p, bias=bool(bias),
                              padding_mode="reflection" if refl else "zeros")
        assert float(conv.mask.sum()) == float(g[f"conv{n}_mask_sum"])
        assert list(conv.state_dict().keys()) == (["mask", "weight", "bias"] if bias else ["mask", "weight"])
        assert not conv.mask.requires_grad and conv.weight.requires_grad
        assert float((conv.weight.detach() * (1 - conv.mask)).abs().sum()) == 0.0
    for lvl in range(4):
        edges = [tuple(e) for e in golden_topology["levels"][lvl]["edges"]]
        pool = H.SkeletonPool(edges, "mean", 3, last_pool=(lvl == 3))
        un = H.SkeletonUnpool(pool.pooling_list, 3)
        assert np.array_equal(pool.weight.numpy(), g[f"pool{lvl}_w"]) and not pool.weight.requires_grad
        assert np.array_equal(un.weight.numpy(), g[f"unpool{lvl}_w"]) and not un.weight.requires_grad


def _kaiming_linear(model):
    import torch.nn.init as init

    def fn(m):
        if m.__class__.__name__.find("Linear") == 0 and hasattr(m, "weight"):
            init.kaiming_normal_(m.weight.data, a=0, mode="fan_in")
            if m.bias is not None:
                init.constant_(m.bias.data, 0.0)
    holder = torch.nn.Module()
    holder.model = model
    holder.apply(fn)


@pytest.mark.parametrize("tag,hp", [("len64", HP64), ("len8", HP8)])
def test_hmvae_state_dict_and_seeded_init(tag, hp, golden_models):
    """Same keys as the reference state_dict and, under the same seed, the same initial weights (RNG order)."""
    g = golden_models
    torch.manual_seed(0)
    model = TwoHierSAVAEModel(hp, device="cpu")
    _kaiming_linear(model)
    sd = model.state_dict()
    assert sorted(sd.keys()) == g[f"{tag}_keys"].tolist()
    assert len(sd) == int(g[f"{tag}_nkeys"])
    for k, v in sd.items():
        if k.startswith("dec.enc."):
            continue
        np.testing.assert_allclose(_cks(v), g[f"{tag}_init/{k}"], rtol=1e-6, atol=1e-6, err_msg=k)
    n_train = sum(p.numel() for p in model.parameters() if p.requires_grad)
    assert n_train == (13233480 if tag == "len64" else n_train)


def test_trajectory_state_dict_and_seeded_init(golden_models):
    g = golden_models
    torch.manual_seed(0)
    model = TrajectoryModel(dict(HPT), device="cpu")
    _kaiming_linear(model)
    for k, v in model.state_dict().items():
        np.testing.assert_allclose(_cks(v), g[f"traj_init/{k}"], rtol=1e-6, atol=1e-6, err_msg=k)


def test_cpu_tensor_is_an_error_not_a_fallback():
    nb = H.find_neighbor(H.get_edges([-1, 0, 0, 1]), 2)
    fk = H.ForwardKinematicsLayer(device="cpu")
    with pytest.raises(Exception):
        fk(torch.zeros(2, 24, 3, 3))
    with pytest.raises(Exception):
        H.rotation_matrix_from_ortho6d(torch.zeros(2, 6))


def test_optimizer_state_is_torch_adam_format_round_trip():
    """optimizer.pt of the reference is torch.optim.Adam.state_dict() (trainer_motion_vae.py:29-31, 112-113, 121-126): the
    converters must (a) read it, (b) write something torch.optim.Adam.load_state_dict accepts, and continuing from the
    round-tripped state must equal an uninterrupted run."""
    import torch

    from hm_vae_b200.optim_state import from_torch_adam, to_torch_adam

    gen = torch.Generator().manual_seed(0)
    shapes = [(6, 4, 3), (6,), (5, 7), (2,)]
    init = [torch.randn(*s, generator=gen) for s in shapes]
    grads = [[torch.randn(*s, generator=gen) for s in shapes] for _ in range(6)]

    def run(params, opt, steps):
        for gs in steps:
            for i, (p, g) in enumerate(zip(params, gs)):
                p.grad = None if i == 3 else g.clone()          # parameter 3 never receives a gradient (reference D9)
            opt.step()

    pa = [torch.nn.Parameter(t.clone()) for t in init]
    oa = torch.optim.Adam(pa, lr=1e-3, weight_decay=1e-4)
    run(pa, oa, grads)                                          # uninterrupted
    pb = [torch.nn.Parameter(t.clone()) for t in init]
    ob = torch.optim.Adam(pb, lr=1e-3, weight_decay=1e-4)
    run(pb, ob, grads[:3])
    step, lr, m, v, live = from_torch_adam(ob.state_dict(), len(pb))          # (a) the reference's file
    assert step == 3 and lr == 1e-3 and live == {0, 1, 2} and m[3] is None
    ours = to_torch_adam(step, lr, (0.9, 0.999), 1e-8, 1e-4, [x if x is not None else torch.zeros(2) for x in m],
                         [x if x is not None else torch.zeros(2) for x in v], live=live)
    assert set(ours) == {"state", "param_groups"} and set(ours["state"]) == {0, 1, 2}
    assert set(ours["param_groups"][0]) >= set(ob.state_dict()["param_groups"][0])
    pc = [torch.nn.Parameter(p.detach().clone()) for p in pb]
    oc = torch.optim.Adam(pc, lr=1e-3, weight_decay=1e-4)
    oc.load_state_dict(ours)                                    # (b)
    run(pc, oc, grads[3:])
    for a, c in zip(pa, pc):
        assert torch.equal(a, c)
    # round-1 layout of this package is still readable
    old = dict(step=3, lr=1e-3, exp_avg=[torch.zeros(*s) for s in shapes], exp_avg_sq=[torch.zeros(*s) for s in shapes])
    assert from_torch_adam(old, 4)[0] == 3
