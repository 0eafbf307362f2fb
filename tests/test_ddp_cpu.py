"""world_size-2 gloo checks of the data-parallel host logic (bucketing, hook-driven all-reduce, unused parameters)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn

from hm_vae_b200.ddp import BucketedAllReduce, broadcast_parameters, make_buckets


class Tiny(nn.Module):
    def __init__(self):
        super().__init__()
        self.a = nn.Linear(8, 16)
        self.unused = nn.Linear(4, 4)      # never receives a gradient (reference D9: latent heads 1, 2)
        self.b = nn.Linear(16, 3)

    def forward(self, x):
        return self.b(torch.relu(self.a(x)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(100 + rank)            # different initial weights per rank ...
        model = Tiny()
        broadcast_parameters(model)              # ... made identical by the one-off broadcast
        ref = [p.detach().clone() for p in model.parameters()]
        sync = BucketedAllReduce(model, n_buckets=2)
        assert sync.world == world
        torch.manual_seed(7 + rank)              # rank-local shard of the batch
        x = torch.randn(5, 8)
        sync.begin()
        model(x).pow(2).mean().backward()
        sync.finish()
        # expected: sum over ranks of the local gradients (the optimiser divides by world)
        local = Tiny()
        with torch.no_grad():
            for p, r in zip(local.parameters(), ref):
                p.copy_(r)
        total = [torch.zeros_like(p) for p in local.parameters()]
        for r in range(world):
            torch.manual_seed(7 + r)
            xr = torch.randn(5, 8)
            local.zero_grad()
            local(xr).pow(2).mean().backward()
            for t, p in zip(total, local.parameters()):
                if p.grad is not None:
                    t += p.grad
        ok = True
        for (n, p), t in zip(model.named_parameters(), total):
            if n.startswith("unused"):
                ok &= p.grad is None
            else:
                ok &= torch.allclose(p.grad, t, rtol=1e-5, atol=1e-6)
        w0 = [p.detach().clone() for p in model.parameters()]
        gathered = [None] * world
        dist.all_gather_object(gathered, [w.tolist() for w in w0])
        ok &= gathered[0] == gathered[1]
        out[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_bucketed_allreduce_gloo_world2():
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert dict(out) == {0: True, 1: True}


def test_bucket_order_is_reverse_of_registration():
    m = Tiny()
    buckets = make_buckets(list(m.named_parameters()), n_buckets=3)
    names = [n for b in buckets for n, _ in b]
    assert names == [n for n, _ in reversed(list(m.named_parameters()))]
    assert sum(len(b) for b in buckets) == 6 and len(buckets) <= 3


# ------------------------------------------------------------------------------------------------ fused DP optimiser (host logic)
def test_arena_layout_and_static_ownership():
    from hm_vae_b200.dp_fused import ALIGN, arena_layout, balanced_bounds, clip_ranges, merge_ranges

    numels = [288 * 144 * 15, 288, 24 * 384, 24, 3, 1008, 7]
    offs, total = arena_layout(numels)
    assert all(o % ALIGN == 0 for o in offs) and total % ALIGN == 0
    assert all(offs[i + 1] >= offs[i] + numels[i] for i in range(len(numels) - 1)) and total >= offs[-1] + numels[-1]
    every = [(offs[i], offs[i] + (numels[i] + ALIGN - 1) // ALIGN * ALIGN) for i in range(len(numels))]
    live = [every[i] for i in (0, 1, 4, 6)]                                     # params 2, 3, 5 got no gradient this step
    assert merge_ranges([(8, 12), (0, 4), (4, 8), (20, 24)]) == [(0, 12), (20, 24)]
    for world in (1, 2, 3, 8):
        bounds = balanced_bounds(every, world, total)                            # static: from ALL parameters, not the step's live set
        assert bounds[0] == 0 and bounds[-1] == total and all(bounds[r] <= bounds[r + 1] for r in range(world))
        covered = []
        for r in range(world):
            own = clip_ranges(live, bounds[r], bounds[r + 1])
            assert all(b % ALIGN == 0 and e % ALIGN == 0 and bounds[r] <= b < e <= bounds[r + 1] for b, e in own)
            covered += own
        assert merge_ranges(covered) == merge_ranges(live)                       # union == live ...
        assert sum(e - b for b, e in covered) == sum(e - b for b, e in merge_ranges(live))   # ... and disjoint


def test_mask_aware_unit_tables():
    """Host logic of the mask-aware fused step (hmvae_dp_adam_step_units): live ranges of a SkeletonConv mask, the ranks'
    balanced static shares and the unit tables -- over all ranks the units cover every unmasked element exactly once and
    (apart from the <= 3 elements of outward rounding per run) nothing else."""
    import numpy as np

    from hm_vae_b200.dp_fused import ALIGN, UNIT, arena_layout, balanced_bounds, clip_ranges, cut_units, mask_live_ranges
    from hm_vae_b200.skeleton import SkeletonConv, find_neighbor, get_edges
    import os as _os
    d = np.load(_os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "hm_vae_b200", "data", "smpl24.npz"))
    parents = d["parents"].tolist()
    edges = get_edges(parents)
    nb = find_neighbor(edges, 2)
    conv = SkeletonConv(nb, 6 * len(nb), 12 * len(nb), 15, len(nb), padding=7, padding_mode='reflection')
    mask = conv.mask.detach().numpy()
    numels = [mask.size, 12 * len(nb), 7, 24 * 384]
    offs, total = arena_layout(numels)
    plive = [mask_live_ranges(mask, offs[0])] + [[(o, o + (n + ALIGN - 1) // ALIGN * ALIGN)] for o, n in zip(offs[1:], numels[1:])]
    flat = mask.reshape(-1) != 0
    cover = np.zeros(total, dtype=np.int32)
    for b, e in plive[0]:
        assert b % ALIGN == 0 and e % ALIGN == 0
        cover[b:e] += 1
    assert cover.max() == 1 and (cover[offs[0]:offs[0] + mask.size][flat] == 1).all()
    runs = len(plive[0])
    assert cover.sum() - flat.sum() <= 2 * (ALIGN - 1) * runs and cover.sum() < 0.9 * mask.size      # most masked entries are skipped
    assert mask_live_ranges(np.zeros((4, 4)), 0) == [] and mask_live_ranges(np.ones((2, 6)), 8) == [(8, 20)]
    all_live = [r for rs in plive for r in rs]
    n_live = sum(e - b for b, e in all_live)
    for world in (1, 2, 3, 8):
        bounds = balanced_bounds(all_live, world, total)
        assert bounds[0] == 0 and bounds[-1] == total and all(bounds[i] <= bounds[i + 1] and bounds[i] % ALIGN == 0 for i in range(world))
        seen = np.zeros(total, dtype=np.int32)
        for r in range(world):
            own = clip_ranges(all_live, bounds[r], bounds[r + 1])
            assert abs(sum(e - b for b, e in own) - n_live / world) <= ALIGN             # balanced on LIVE elements
            table = cut_units(own)
            assert table.dtype == np.int32 and table.shape[1] == 2 and (table[:, 1] >= 1).all() and (table[:, 1] <= UNIT).all()
            for f, n in table:
                seen[ALIGN * f:ALIGN * (f + n)] += 1
        want = np.zeros(total, dtype=np.int32)
        for b, e in all_live:
            want[b:e] = 1
        assert (seen == want).all()
    assert cut_units([]).shape == (0, 2)


def _dp_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from hm_vae_b200.dp_fused import arena_layout

        # the reduce-scatter / sharded-Adam / all-gather dataflow of hmvae_dp_adam_step, emulated with gloo collectives on the
        # host: every rank updates only its owned ranges from the SUM of all ranks' gradients, then the ranks exchange their
        # shares; the result must equal a plain Adam step on the averaged gradient.
        torch.manual_seed(0)
        shapes = [(6, 5, 3), (6,), (4, 10), (3,)]
        params = [torch.randn(*s) for s in shapes]
        offs, total = arena_layout([p.numel() for p in params])
        p_arena, g_arena = torch.zeros(total), torch.zeros(total)
        for p, o in zip(params, offs):
            p_arena[o:o + p.numel()] = p.reshape(-1)
        torch.manual_seed(10 + rank)
        for p, o in zip(params, offs):
            g_arena[o:o + p.numel()] = torch.randn(p.numel())
        # tensor 0 is a masked SkeletonConv-style weight: dead entries hold zero values and zero gradients and are in no work unit
        from hm_vae_b200.dp_fused import balanced_bounds, clip_ranges, cut_units, mask_live_ranges
        mask0 = torch.zeros(shapes[0])
        mask0[:3, :2] = 1
        mask0[3:, 2:] = 1
        dead0 = (mask0 == 0).reshape(-1)
        p_arena[offs[0]:offs[0] + mask0.numel()][dead0] = 0
        g_arena[offs[0]:offs[0] + mask0.numel()][dead0] = 0
        params[0] = params[0] * mask0
        plive = [mask_live_ranges(mask0.numpy(), offs[0])] + [[(o, o + (p.numel() + 3) // 4 * 4)] for p, o in zip(params[1:], offs[1:])]
        live = [r for rs in plive for r in rs]
        bounds = balanced_bounds(live, world, total)                 # static shares, balanced on the live elements
        gsum = g_arena.clone()
        dist.all_reduce(gsum)                                        # what the peer loads add up to
        lr, b1, b2, eps, wd = 1e-2, 0.9, 0.999, 1e-8, 1e-4
        new = torch.zeros(total)
        touched = torch.zeros(total)
        for first4, n4 in cut_units(clip_ranges(live, bounds[rank], bounds[rank + 1])).tolist():      # this rank's unit table
            b, e = 4 * first4, 4 * (first4 + n4)
            g = gsum[b:e] / world + wd * p_arena[b:e]
            m, v = (1 - b1) * g, (1 - b2) * g * g
            new[b:e] = p_arena[b:e] - lr / (1 - b1) * m / (v.sqrt() / (1 - b2) ** 0.5 + eps)
            touched[b:e] += 1
        dist.all_reduce(new)                                         # shares are disjoint: the sum is the all-gather
        dist.all_reduce(touched)
        assert float(touched.max()) == 1.0                           # every element stepped by at most one rank ...
        t0 = touched[offs[0]:offs[0] + mask0.numel()]
        assert bool((t0[~dead0] == 1).all()) and float(t0[dead0].sum()) < 0.5 * float(dead0.sum())    # ... live ones exactly once
        ref_p = [p.clone().requires_grad_(True) for p in params]
        opt = torch.optim.Adam(ref_p, lr=lr, weight_decay=wd)
        for p, o in zip(ref_p, offs):
            p.grad = (gsum[o:o + p.numel()] / world).view(p.shape).clone()
        opt.step()
        ok = all(torch.allclose(new[o:o + p.numel()].view(p.shape), p.detach(), rtol=1e-5, atol=1e-7) for p, o in zip(ref_p, offs))
        out[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_fused_dp_dataflow_gloo_world2():
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_dp_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert dict(out) == {0: True, 1: True}
