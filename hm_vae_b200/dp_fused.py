"""Data-parallel Adam fused with its collective over NVLink peer memory (``hmvae_dp_adam_step``, csrc/dp.cu).

Replaces ``torch.nn.DataParallel`` + ``torch.optim.Adam`` of the reference (train_motion_vae.py:49-53,
trainer_motion_vae.py:29-31, 92-93).  Parameters and gradients of all trainable tensors live in two flat, symmetric
arenas per rank (same offset = same element on every rank), mapped into every peer's address space
(``torch.distributed._symmetric_memory``, or CUDA IPC through the C ABI when that is unavailable).  The weight / bias
gradient kernels write straight into the gradient arena (``ops.grad_buffer``), so one kernel per step does
reduce-scatter -> sharded Adam -> all-gather; there is no NCCL call on the step path and the whole step, collective
included, is one CUDA graph.

The pure-Python helpers (``arena_layout``, ``merge_ranges``, ``mask_live_ranges``, ``balanced_bounds``, ``clip_ranges``,
``cut_units``) are covered by the CPU tests (gloo, world 2).
"""
import ctypes
import os

import torch
import torch.distributed as dist

from . import _lib

ALIGN = 4          # elements: every tensor starts on a 16-byte boundary; ranges are multiples of 4 (float4 path)


def arena_layout(numels):
    """[numel] -> ([offset], total) with every offset (and the total) a multiple of ALIGN."""
    offs, cur = [], 0
    for n in numels:
        offs.append(cur)
        cur += (int(n) + ALIGN - 1) // ALIGN * ALIGN
    return offs, cur


def merge_ranges(ranges):
    """Sorted, disjoint, adjacent-merged [begin, end) list."""
    out = []
    for b, e in sorted(ranges):
        if e <= b:
            continue
        if out and b <= out[-1][1]:
            out[-1][1] = max(out[-1][1], e)
        else:
            out.append([b, e])
    return [tuple(r) for r in out]


UNIT = 32          # float4 per work unit of hmvae_dp_adam_step_units (one per lane of a warp)


def mask_live_ranges(mask, offset):
    """Arena ranges [begin, end) (multiples of ALIGN, merged) covering the non-zero entries of a flattened 0/1 ``mask`` that
    starts at arena element ``offset`` (itself a multiple of ALIGN).  Rounding outwards may include a few masked elements:
    they hold zero parameters and zero gradients, on which Adam is the identity."""
    import numpy as np

    flat = np.asarray(mask).reshape(-1) != 0
    if flat.size == 0 or not flat.any():
        return []
    edges = np.flatnonzero(np.diff(np.concatenate(([False], flat, [False])).astype(np.int8)))
    begs, ends = edges[0::2], edges[1::2]
    begs = begs // ALIGN * ALIGN
    ends = (ends + ALIGN - 1) // ALIGN * ALIGN
    return merge_ranges([(int(offset + b), int(offset + e)) for b, e in zip(begs, ends)])


def balanced_bounds(live, world, total):
    """W + 1 arena positions that split the LIVE elements of ``live`` (merged ranges) evenly: rank r owns arena elements
    [bounds[r], bounds[r+1]).  Static (it depends on the masks only), so a rank's slice of the Adam moments never migrates."""
    live = merge_ranges(live)
    n4 = sum((e - b) // ALIGN for b, e in live)
    bounds = [0]
    for r in range(1, world):
        want, seen, pos = (n4 * r) // world, 0, total
        for b, e in live:
            k = (e - b) // ALIGN
            if seen + k >= want:
                pos = b + (want - seen) * ALIGN
                break
            seen += k
        bounds.append(max(pos, bounds[-1]))
    bounds.append(total)
    return bounds


def clip_ranges(ranges, lo, hi):
    out = []
    for b, e in merge_ranges(ranges):
        s, t = max(b, lo), min(e, hi)
        if t > s:
            out.append((s, t))
    return out


def cut_units(ranges):
    """Merged element ranges -> int32 [n, 2] table {first float4, number of float4 <= UNIT} for hmvae_dp_adam_step_units."""
    import numpy as np

    firsts, counts = [], []
    for b, e in ranges:
        b4, n4 = b // ALIGN, (e - b) // ALIGN
        starts = np.arange(0, n4, UNIT, dtype=np.int64)
        firsts.append(b4 + starts)
        counts.append(np.minimum(UNIT, n4 - starts))
    if not firsts:
        return np.zeros((0, 2), dtype=np.int32)
    out = np.stack([np.concatenate(firsts), np.concatenate(counts)], axis=1)
    assert out[:, 0].max(initial=0) < 2 ** 31
    return np.ascontiguousarray(out.astype(np.int32))


class _RawCudaBuffer:
    """A device allocation owned by the C library, exposed to torch through __cuda_array_interface__."""

    def __init__(self, ptr, numel, typestr):
        self.ptr, self.numel = ptr, numel
        self.__cuda_array_interface__ = {"shape": (numel,), "typestr": typestr, "data": (ptr, False), "version": 2}


class PeerArenas:
    """grad / param arenas (float32[numel]) and a flag pad (uint32) mapped on every rank of ``group``."""

    def __init__(self, numel, device, group=None, prefer="symm"):
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.numel, self.device, self.group = numel, device, group
        self.backend = None
        self.mc_grad = self.mc_param = 0
        self._keep = []
        if self.world == 1:
            self.grad = torch.zeros(numel, device=device, dtype=torch.float32)
            self.param = torch.zeros(numel, device=device, dtype=torch.float32)
            self.flags = torch.zeros(64, device=device, dtype=torch.int32)
            self.grad_ptrs, self.param_ptrs, self.flag_ptrs = [self.grad.data_ptr()], [self.param.data_ptr()], [self.flags.data_ptr()]
            self.backend = "local"
            return
        err = None
        if prefer == "symm":
            try:
                self._init_symm()
                return
            except Exception as exc:         # noqa: BLE001 -- fall through to CUDA IPC
                err = exc
        try:
            self._init_ipc()
        except Exception as exc2:            # noqa: BLE001
            raise _lib.HmvaeError("peer memory unavailable: symmetric memory: %r; CUDA IPC: %r" % (err, exc2))

    def _init_symm(self):
        import torch.distributed._symmetric_memory as symm_mem

        group = self.group if self.group is not None else dist.group.WORLD
        bufs, hdls = [], []
        for n, dt in ((self.numel, torch.float32), (self.numel, torch.float32), (64, torch.int32)):
            t = symm_mem.empty(n, dtype=dt, device=self.device)
            t.zero_()
            hdls.append(symm_mem.rendezvous(t, group))
            bufs.append(t)
        torch.cuda.synchronize()
        dist.barrier(group=self.group)
        self.grad, self.param, self.flags = bufs
        self.grad_ptrs, self.param_ptrs, self.flag_ptrs = [list(h.buffer_ptrs) for h in hdls]
        self._keep = hdls
        self.backend = "symmetric_memory"
        # NVSwitch multicast (NVLS) mappings, when the fabric offers them
        try:
            mg, mp_ = int(hdls[0].multicast_ptr), int(hdls[1].multicast_ptr)
        except Exception:       # noqa: BLE001
            mg = mp_ = 0
        # measured (B200, NVSwitch): ms/step unicast vs multicast = 0.955 / 1.007 (2 GPUs), 1.014 / 0.979 (4), 1.080 / 1.020 (8)
        want = os.environ.get("HMVAE_DP_MULTICAST", "auto")
        use = (self.world >= 3) if want == "auto" else (want != "0")
        if mg and mp_ and use:
            self.mc_grad, self.mc_param = mg, mp_
            self.backend = "symmetric_memory+nvls_multicast"

    def _init_ipc(self):
        lib = _lib.lib
        ptrs = []
        for nbytes in (self.numel * 4, self.numel * 4, 64 * 4):
            p = ctypes.c_void_p()
            _lib.check(lib.hmvae_ipc_alloc(nbytes, ctypes.byref(p)), "ipc_alloc")
            ptrs.append(p.value)
        handles = []
        for p in ptrs:
            h = (ctypes.c_ubyte * 64)()
            _lib.check(lib.hmvae_ipc_get_handle(ctypes.c_void_p(p), h), "ipc_get_handle")
            handles.append(bytes(h))
        gathered = [None] * self.world
        dist.all_gather_object(gathered, handles, group=self.group)
        tables = [[], [], []]
        for q in range(self.world):
            for k in range(3):
                if q == self.rank:
                    tables[k].append(ptrs[k])
                else:
                    out = ctypes.c_void_p()
                    hb = (ctypes.c_ubyte * 64).from_buffer_copy(gathered[q][k])
                    _lib.check(lib.hmvae_ipc_open_handle(hb, ctypes.byref(out)), "ipc_open_handle")
                    tables[k].append(out.value)
        self.grad_ptrs, self.param_ptrs, self.flag_ptrs = tables
        self._raw = [_RawCudaBuffer(ptrs[0], self.numel, "<f4"), _RawCudaBuffer(ptrs[1], self.numel, "<f4"), _RawCudaBuffer(ptrs[2], 64, "<i4")]
        self.grad = torch.as_tensor(self._raw[0], device=self.device)
        self.param = torch.as_tensor(self._raw[1], device=self.device)
        self.flags = torch.as_tensor(self._raw[2], device=self.device)
        torch.cuda.synchronize()
        dist.barrier(group=self.group)
        self.backend = "cuda_ipc"


class FusedDataParallelAdam:
    """torch.optim.Adam(lr, betas, eps, weight_decay) over flat peer-mapped arenas; ``step_dyn`` is the fused
    reduce-scatter + Adam + all-gather kernel.  Same host interface as ``ops.FusedAdam`` (advance / step_dyn / zero_grad)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, group=None, prefer="symm", masks=None):
        """``masks``: {id(parameter): 0/1 tensor of the parameter's shape} for parameters whose masked entries are structurally
        dead (SkeletonConv.weight / .mask): zero value, zero gradient on every step.  Those entries are left out of the unit
        table, i.e. never read or written by the step.  A parameter whose masked entries are NOT all zero when the
        optimiser is built is treated as fully live (weight decay would move them in the reference)."""
        from . import ops

        self.params = [p for p in params]
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.step_count = 0            # host mirrors of the device clock {Adam steps, scheduler iterations}
        self.sched_iters = 0
        self.gamma, self.step_size = 1.0, 0
        self.param_groups = [dict(lr=lr)]
        self.group = group
        self._live = set()
        dev = self.params[0].device
        self.offsets, self.numel = arena_layout([p.numel() for p in self.params])
        self.arenas = PeerArenas(self.numel, dev, group=group, prefer=prefer)
        self.world, self.rank = self.arenas.world, self.arenas.rank
        # static, mask-aware liveness per parameter and the ranks' (balanced) shares of it
        self._plive, self.masked_elems = [], 0
        use_masks = os.environ.get("HMVAE_DP_MASK_AWARE", "1") != "0"
        for p, off in zip(self.params, self.offsets):
            n = p.numel()
            full = [(off, off + (n + ALIGN - 1) // ALIGN * ALIGN)]
            mk = masks.get(id(p)) if (masks and use_masks) else None
            if mk is not None and tuple(mk.shape) == tuple(p.shape):
                mk = mk.detach().to(p.device)
                if float((p.detach() * (mk == 0)).abs().max()) == 0.0:
                    rs = mask_live_ranges(mk.cpu().numpy(), off)
                    self.masked_elems += n - sum(e - b for b, e in rs)
                    self._plive.append(rs)
                    continue
            self._plive.append(full)
        self.bounds = balanced_bounds([r for rs in self._plive for r in rs], self.world, self.numel)
        self.m = torch.zeros(self.numel, device=dev, dtype=torch.float32)
        self.v = torch.zeros(self.numel, device=dev, dtype=torch.float32)
        self.state = torch.zeros(4, device=dev, dtype=torch.int32)
        self._clock = torch.zeros(2, dtype=torch.int32, device=dev)
        self._dyn_dev = torch.zeros(2, dtype=torch.float32, device=dev)
        # parameters move into the arena (identical values on every rank are the caller's job: broadcast first)
        with torch.no_grad():
            for p, off in zip(self.params, self.offsets):
                view = self.arenas.param[off:off + p.numel()].view(p.shape)
                view.copy_(p.data)
                p.data = view
                p.grad = None
        # ... and the gradient kernels write into the gradient arena
        for p, off in zip(self.params, self.offsets):
            ops.register_grad_buffer(p, self.arenas.grad, off)
        peers = _lib.DpPeers()
        peers.world, peers.rank = self.world, self.rank
        for q in range(self.world):
            peers.grad[q] = self.arenas.grad_ptrs[q]
            peers.param[q] = self.arenas.param_ptrs[q]
            peers.flags[q] = self.arenas.flag_ptrs[q]
        peers.mc_grad = self.arenas.mc_grad or None
        peers.mc_param = self.arenas.mc_param or None
        self._peers = peers
        self._range_cache = {}
        torch.cuda.synchronize()
        if self.world > 1:
            dist.barrier(group=group)

    def close(self):
        from . import ops

        ops.unregister_grad_buffers(self.arenas.grad)

    def __del__(self):
        try:
            self.close()
        except Exception:       # noqa: BLE001 -- interpreter shutdown
            pass

    # ---- same surface as ops.FusedAdam
    def zero_grad(self, set_to_none=True):
        for p in self.params:
            p.grad = None

    def set_schedule(self, gamma, step_size):
        """StepLR(step_size, gamma) evaluated on the device (hmvae_opt_clock_tick); step_size <= 0 = constant."""
        self.gamma, self.step_size = float(gamma), int(step_size)

    def current_lr(self):
        if self.step_size > 0:
            return self.lr * self.gamma ** (self.sched_iters // self.step_size)
        return self.lr

    def set_clock(self, step, iterations):
        """Synchronous (checkpoint resume): Adam step count and scheduler position."""
        self.step_count, self.sched_iters = int(step), int(iterations)
        self._clock.copy_(torch.tensor([self.step_count, self.sched_iters], dtype=torch.int32))
        self.param_groups[0]["lr"] = self.current_lr()

    def advance(self, lr=None):
        """Host side of one step: only the mirrors move -- the step counter and the schedule position live on the device and
        tick inside the (possibly replayed) step, so a host that runs ahead cannot change the scalars of a queued step."""
        self.param_groups[0]["lr"] = self.current_lr()
        self.step_count += 1
        self.sched_iters += 1

    def _live_ranges(self, select, written=()):
        """Live = parameters (with index in ``select``) that received a gradient this step.  Gradients that autograd did not
        place in the arena (it clones a gradient it cannot steal) are copied in.  ``written``: ids of parameters for which the
        caller vouches that their gradient kernels have been issued and wrote into the arena (a call from INSIDE a backward
        node, before autograd has assigned ``.grad``)."""
        key = []
        for i in select:
            p, off = self.params[i], self.offsets[i]
            if p.grad is None and id(p) not in written:
                continue
            n = p.numel()
            if p.grad is not None and p.grad.data_ptr() != self.arenas.grad.data_ptr() + 4 * off:
                self.arenas.grad[off:off + n].view(p.shape).copy_(p.grad)
            key.append(i)
            self._live.add(i)
        key = tuple(key)
        if key not in self._range_cache:
            if torch.cuda.is_current_stream_capturing():
                raise _lib.HmvaeError("fused data-parallel step: run one eager step with this set of live parameters before "
                                      "capturing it in a CUDA graph (the unit table is uploaded on first use)")
            lo, hi = self.bounds[self.rank], self.bounds[self.rank + 1]
            own = clip_ranges([r for i in key for r in self._plive[i]], lo, hi)
            table = cut_units(own)
            dev_table = torch.from_numpy(table).to(self.arenas.grad.device) if len(table) else \
                torch.zeros((1, 2), dtype=torch.int32, device=self.arenas.grad.device)
            self._range_cache[key] = (dev_table, len(table), key)
        return self._range_cache[key]

    def _launch(self, select, grad_scale, max_ctas=0, in_flight=0, written=()):
        from . import ops

        table, n, live = self._live_ranges(select, written)
        if getattr(self, "_tick_due", False):
            _lib.check(_lib.lib.hmvae_opt_clock_tick(self._clock.data_ptr(), self.lr, self.gamma, self.step_size, self.betas[0],
                                                     self.betas[1], _lib.ptr(self._dyn_dev), ops.stream()), "opt_clock_tick")
            self._tick_due = False
        for i in live:
            torch.autograd.graph.increment_version(self.params[i])
        scale = (1.0 / self.world) if grad_scale is None else grad_scale
        _lib.check(_lib.lib.hmvae_dp_adam_step_units(ctypes.byref(self._peers), _lib.ptr(self.m), _lib.ptr(self.v),
                                                     table.data_ptr(), n, _lib.ptr(self._dyn_dev), self.betas[0], self.betas[1],
                                                     self.eps, self.weight_decay, scale, self.state.data_ptr(), int(max_ctas),
                                                     int(in_flight), ops.stream()), "dp_adam_step_units")

    def begin_step(self):
        """Start of a device step (capturable): the device clock ticks and refreshes the step-dependent scalars; nothing has
        been stepped yet."""
        from . import ops

        # The tick itself is issued lazily, on the optimiser's side stream right before the step's first optimiser kernel
        # (_launch): only those kernels read its scalars, and as a dependency-free node at the head of a replayed graph it would
        # be dispatched ahead of the forward pass's first kernels (tools/timeline.py).
        self._tick_due = True
        self._stepped = set()
        self._begun = True

    def step_partial(self, param_ids, grad_scale=None, written=()):
        """Steps ONLY the given parameters, now, on a side stream: called in the middle of the backward pass as soon as their
        gradients are final (the decoder's, while the encoder's backward still runs; the deepest encoder level's, while the
        shallower levels' backward still runs), so that the collective + optimiser work of each bucket hides under the rest of
        backward -- buckets in reverse-layer order.  Every rank must make the same sequence of calls.  ``written``: see
        ``_live_ranges``."""
        from . import ops

        if not getattr(self, "_begun", False):
            self.begin_step()
        select = [i for i, p in enumerate(self.params) if id(p) in param_ids and i not in self._stepped]
        if not select:
            return
        if getattr(self, "_opt_stream", None) is None:
            self._opt_stream = torch.cuda.Stream()
        main = torch.cuda.current_stream()
        self._opt_stream.wait_stream(main)
        for s in ops._overlap.get("used", ()) or ():           # weight gradients are produced on the side stream(s)
            self._opt_stream.wait_stream(s)
        with torch.cuda.stream(self._opt_stream):
            # across ranks: a small grid that hides the NVLink latency with loads in flight instead of threads, so that the
            # bucket does not take the SMs of the backward pass it runs under (8 GPUs: 32 CTAs x 4 beat 64 x 2 and the full grid)
            multi = self.world > 1
            self._launch(select, grad_scale, max_ctas=int(os.environ.get("HMVAE_DP_PARTIAL_CTAS", "32" if multi else "296")),
                         in_flight=int(os.environ.get("HMVAE_DP_PARTIAL_IN_FLIGHT", "4" if multi else "0")), written=written)
        self._stepped.update(select)
        self._partial_pending = True

    def step_dyn(self, grad_scale=None):
        """Device side (capturable): everything not stepped by ``step_partial`` in this step.  ``grad_scale`` defaults to
        1/world (mean of the per-rank gradients)."""
        if not getattr(self, "_begun", False):
            self.begin_step()
        if getattr(self, "_partial_pending", False):
            torch.cuda.current_stream().wait_stream(self._opt_stream)     # also orders the two kernels' epochs / flags
            self._partial_pending = False
        self._launch([i for i in range(len(self.params)) if i not in self._stepped], grad_scale)
        self._begun = False

    def step(self, grad_scale=None, lr=None):
        self.advance(lr)
        self.step_dyn(grad_scale)

    def timed_out(self):
        """True if a flag barrier ever timed out (a peer died or fell > HMVAE_DP_TIMEOUT_S behind): the kernel then SKIPPED its
        update (parameters stay consistent but stale).  Synchronises with the device."""
        return bool(int(self.state[2].item()))

    def check_health(self):
        if self.timed_out():
            raise _lib.HmvaeError("fused data-parallel step: a peer did not reach the gradient barrier within HMVAE_DP_TIMEOUT_S; "
                                  "the update was skipped on this rank -- the ranks are out of step, abort the job")

    # ---- checkpoints: every rank only maintains the moments of its static share of the arena
    def _full_moments(self):
        """COLLECTIVE over ``self.group`` when world > 1: every rank must call it."""
        lo, hi = self.bounds[self.rank], self.bounds[self.rank + 1]
        m, v = torch.zeros_like(self.m), torch.zeros_like(self.v)
        m[lo:hi] = self.m[lo:hi]
        v[lo:hi] = self.v[lo:hi]
        if self.world > 1:
            dist.all_reduce(m, group=self.group)
            dist.all_reduce(v, group=self.group)
        return m, v

    def state_dict(self):
        """torch.optim.Adam.state_dict() layout (the reference's optimizer.pt).  COLLECTIVE when world > 1 (the moments are
        sharded): every rank must call it; all ranks get the full dict."""
        from .optim_state import to_torch_adam

        m, v = self._full_moments()
        return to_torch_adam(self.step_count, self.current_lr(), self.betas, self.eps, self.weight_decay,
                             [m[o:o + p.numel()].view(p.shape) for p, o in zip(self.params, self.offsets)],
                             [v[o:o + p.numel()].view(p.shape) for p, o in zip(self.params, self.offsets)],
                             live=self._live, initial_lr=self.lr)

    def load_state_dict(self, sd):
        from .optim_state import from_torch_adam

        step, lr, ms, vs, live = from_torch_adam(sd, len(self.params))
        self.m.zero_()
        self.v.zero_()
        for p, o, a, b in zip(self.params, self.offsets, ms, vs):
            if a is not None:
                self.m[o:o + p.numel()].copy_(a.reshape(-1))
                self.v[o:o + p.numel()].copy_(b.reshape(-1))
        self._live = set(live)
        self.set_clock(step, self.sched_iters)
