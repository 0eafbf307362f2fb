"""Batch-sharded data parallelism: one process per GPU, NCCL all-reduce of gradients over NVLink 5 / NVSwitch.

Replaces ``torch.nn.DataParallel`` (train_motion_vae.py:49-53), which re-broadcasts all 111.8 MB of parameters and
frozen buffers every step and reduces onto GPU 0.  Here weights are broadcast once; each step the gradient tensors
are all-reduced in place, grouped into a few buckets in reverse-layer order, each bucket launched on a side stream
as soon as its last gradient has been produced by the backward pass (so dec/enc wgrad keeps running underneath).
Parameters that receive no gradient in the reference (D9: latent heads 1,2) are simply never in a bucket; which
ones are live is a static property of (model, iterations < iteration_interval), so every rank agrees.
"""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """torchrun environment -> (rank, world, local_rank).  Single process when WORLD_SIZE is unset."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, world, local


def broadcast_parameters(module, src=0):
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src)


def make_buckets(named_params, n_buckets=4):
    """Splits parameters (given in forward/registration order) into ~equal-byte buckets in REVERSE order, i.e. the
    order in which the backward pass produces their gradients."""
    params = [(n, p) for n, p in named_params if p.requires_grad]
    params.reverse()
    total = sum(p.numel() for _, p in params)
    target = max(total // max(n_buckets, 1), 1)
    buckets, cur, acc = [], [], 0
    for n, p in params:
        cur.append((n, p))
        acc += p.numel()
        if acc >= target and len(buckets) < n_buckets - 1:
            buckets.append(cur)
            cur, acc = [], 0
    if cur:
        buckets.append(cur)
    return buckets


class BucketedAllReduce:
    """Overlapped gradient all-reduce.  Usage per step:  sync.begin(); <backward>; sync.finish()."""
    fused = False

    def __init__(self, module, n_buckets=4, live=None, group=None):
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.group = group
        named = [(n, p) for n, p in module.named_parameters() if not n.startswith("dec.enc.")]
        if live is not None:
            named = [(n, p) for n, p in named if n in live]
        self.buckets = make_buckets(named, n_buckets)
        self._bucket_of = {}
        self._pending = []
        self._hooks = []
        self.comm_stream = torch.cuda.Stream() if (self.world > 1 and torch.cuda.is_available()) else None
        for bi, bucket in enumerate(self.buckets):
            for _, p in bucket:
                self._bucket_of[id(p)] = bi
                if self.world > 1:
                    self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))
        self.launched = 0
        self.enabled = True       # False: hooks are inert (used while a CUDA graph of fwd+bwd is captured / replayed)

    def allreduce_now(self):
        """All buckets, in backward order, on the current stream (no overlap): the step's fwd+bwd ran as a CUDA graph."""
        if self.world <= 1:
            return
        for bucket in self.buckets:
            grads = [p.grad for _, p in bucket if p.grad is not None]
            if not grads:
                continue
            with dist._coalescing_manager(group=self.group, device=grads[0].device, async_ops=False):
                for g in grads:
                    dist.all_reduce(g, group=self.group)

    def begin(self):
        self._pending = [len(b) for b in self.buckets]
        self.launched = 0

    def _on_grad(self, p):
        if not self.enabled:
            return
        bi = self._bucket_of[id(p)]
        self._pending[bi] -= 1
        if self._pending[bi] == 0:
            self._launch(bi)

    def _launch(self, bi):
        grads = [p.grad for _, p in self.buckets[bi] if p.grad is not None]
        if not grads:
            return
        self.launched += 1
        if self.comm_stream is None:
            for g in grads:
                dist.all_reduce(g, group=self.group)
            return
        self.comm_stream.wait_stream(torch.cuda.current_stream())
        from . import ops
        if ops._overlap["pending"]:
            for s in ops._overlap.get("used", ()) or ():             # weight gradients are produced on the side streams
                self.comm_stream.wait_stream(s)
        with torch.cuda.stream(self.comm_stream):
            with dist._coalescing_manager(group=self.group, device=grads[0].device, async_ops=False):
                for g in grads:
                    dist.all_reduce(g, group=self.group)
            for g in grads:
                g.record_stream(self.comm_stream)

    def finish(self):
        """Flushes buckets whose hooks did not all fire (unused parameters) and joins the comm stream."""
        if self.world <= 1 or not self.enabled:
            return
        for bi, left in enumerate(self._pending):
            if left > 0:
                self._pending[bi] = 0
                self._launch(bi)
        if self.comm_stream is not None:
            torch.cuda.current_stream().wait_stream(self.comm_stream)

    def remove(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []
