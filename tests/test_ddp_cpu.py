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
