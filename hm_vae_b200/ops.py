"""torch.autograd bridges over the C ABI (include/hmvae_b200.h).

Every op here launches hand-written sm_100a kernels from libhmvae_b200.so on the current CUDA stream, on
caller-owned fp32 tensors; outputs come from PyTorch's caching allocator.  There is no CPU path.
"""
import contextlib
import ctypes
import os

import torch
from torch.autograd import Function

from . import _lib
from ._lib import check, int_array, lib, ptr, stream

IMPL_AUTO, IMPL_SIMT, IMPL_TC = 0, 1, 2
_conv_impl = IMPL_AUTO


def set_conv_impl(impl):
    """0 = auto, 1 = CUDA-core fp32, 2 = tcgen05 TF32 (parity tests pin one or the other)."""
    global _conv_impl
    _conv_impl = int(impl)


# ------------------------------------------------------------------------------------------------ conv plan
class ConvPlan:
    """Immutable per-layer index tables on the device (neighbour CSR, transpose CSR, block list, unpool map)."""

    def __init__(self, neighbour_list, ci, co, ksize, stride, pad, pad_mode, upsample=False, unpool_src=None,
                 src_joints=None, lrelu=False, out_joint_stride=0, out_chan_offset=0, out_channels_last=False):
        j = len(neighbour_list)
        self.joints, self.ci, self.co, self.ksize, self.stride, self.pad = j, ci, co, ksize, stride, pad
        self.pad_mode = {"constant": 0, "zeros": 0, "reflect": 1, "reflection": 1}[pad_mode]
        self.upsample, self.lrelu = bool(upsample), bool(lrelu)
        self.src_joints = int(src_joints) if unpool_src is not None else j
        self.out_joint_stride = out_joint_stride or co
        self.out_chan_offset = out_chan_offset
        self.out_channels_last = bool(out_channels_last)
        self.has_prologue = self.upsample or unpool_src is not None
        off, idx = [0], []
        for nb in neighbour_list:
            idx.extend(int(k) for k in nb)
            off.append(len(idx))
        desc = _lib.ConvDesc(j, ci, co, ksize, stride, pad, self.pad_mode, int(self.upsample), self.src_joints,
                             int(self.lrelu), self.out_joint_stride, out_chan_offset, int(self.out_channels_last))
        handle = ctypes.c_void_p()
        src = int_array(unpool_src) if unpool_src is not None else None
        check(lib.hmvae_conv_plan_create(ctypes.byref(desc), int_array(off), int_array(idx), src, ctypes.byref(handle)),
              "conv_plan_create")
        self.handle = handle
        self.device = torch.cuda.current_device()

    def t_out(self, t_in):
        return (t_in + 2 * self.pad - self.ksize) // self.stride + 1

    def __del__(self):
        h, self.handle = getattr(self, "handle", None), None
        if h:
            try:
                lib.hmvae_conv_plan_destroy(h)
            except Exception:
                pass


_force_repack = False
_overlap = {"on": False, "side": None, "pending": [], "allowed": True}


def _side_stream():
    """The stream of the weight packing (forward) -- also side stream 0 of the weight gradients."""
    if _overlap["side"] is None:
        _overlap["side"] = torch.cuda.Stream(priority=0)
    return _overlap["side"]


def main_stream_priority():
    """Priority of the stream a training step is captured on (Trainer.enable_cuda_graph).  The forward / data-gradient chain is
    the critical path of a step; weight gradients, weight packing and the optimiser run beside it on default-priority streams.
    With a higher priority (HMVAE_MAIN_PRIORITY, default -1) the block scheduler serves the chain's CTAs first whenever both
    have blocks pending (tools/timeline.py: dgrad kernels took 20-33 us next to the weight gradients, 13-23 us alone)."""
    return int(os.environ.get("HMVAE_MAIN_PRIORITY", "-1"))


def _eps_stream():
    """Stream of the N(0,1) draws of a step: they depend on nothing, but queued behind the weight packing they would come too late,
    and queued before it they delay the first conv (tools/timeline.py)."""
    if _overlap.get("eps") is None:
        _overlap["eps"] = torch.cuda.Stream(priority=0)
    return _overlap["eps"]


def _prestage_stream():
    """Stream of the weight gradients' x-operand staging issued during the forward pass (stack._prestage_x)."""
    if _overlap.get("prestage") is None:
        _overlap["prestage"] = torch.cuda.Stream(priority=0)
    return _overlap["prestage"]


def _wgrad_stream():
    """Weight gradients of consecutive layers go round-robin to HMVAE_WGRAD_STREAMS side streams.  Default 2: with the linked
    stack path the data-gradient chain is short enough that ONE stream of weight gradients (prep -> bias -> tcgen05 kernel per
    layer, ~375 us of sequential kernels) became the longer chain of the backward pass (same box, graph replay: 836 / 807 / 826 us
    per step with 1 / 2 / 3 streams)."""
    n = max(1, int(os.environ.get("HMVAE_WGRAD_STREAMS", "2")))
    pool = _overlap.setdefault("pool", [])
    while len(pool) < n:
        pool.append(_side_stream() if not pool else torch.cuda.Stream(priority=0))
    _overlap["rr"] = (_overlap.get("rr", -1) + 1) % n
    s = pool[_overlap["rr"]]
    _overlap.setdefault("used", set()).add(s)
    return s


def join_wgrad():
    """Makes the current stream wait for the weight gradients issued on the side stream."""
    if _overlap["pending"]:
        for s in _overlap.get("used", ()) or (_overlap["side"],):
            torch.cuda.current_stream().wait_stream(s)
        _overlap.get("used", set()).clear()
        _overlap["pending"].clear()


class wgrad_overlap:
    """Context manager: conv weight gradients of backward passes run inside it go to a side stream; joined on exit."""

    def __enter__(self):
        _overlap["on"] = _overlap["allowed"]        # "allowed" is cleared by bench.py's per-kernel attribution pass
        return self

    def __exit__(self, *exc):
        _overlap["on"] = False
        join_wgrad()
        return False
_wgrad_tc = True          # tensor-core weight gradient where the geometry is supported (False: CUDA-core wgrad)

# ------------------------------------------------------------------------------------------------ gradient arena
# dp_fused.FusedDataParallelAdam keeps all gradients in one flat, peer-mapped buffer: the weight / bias gradient kernels
# then write straight into it (key: the parameter's data pointer, which is stable -- parameters live in an arena too).
_grad_arena = {}


def register_grad_buffer(param, arena, offset):
    _grad_arena[param.data_ptr()] = (arena, int(offset), tuple(param.shape), param.numel())


def unregister_grad_buffers(arena=None):
    """Drops every registration (or only those that point into ``arena``)."""
    if arena is None:
        _grad_arena.clear()
        return
    for k in [k for k, e in _grad_arena.items() if e[0] is arena]:
        del _grad_arena[k]


def grad_buffer(param, zero=False):
    """A fresh tensor for the gradient of ``param``: a view of the gradient arena when the parameter is registered (masked
    conv blocks there are zero forever: nothing ever writes them), else a new allocation."""
    e = _grad_arena.get(param.data_ptr())
    if e is not None and e[2] == tuple(param.shape):
        arena, off, shape, n = e
        return arena[off:off + n].view(shape)
    return torch.zeros_like(param) if zero else torch.empty_like(param)


class PackedWeights:
    """Derived cache of a conv weight for the tcgen05 kernels: tf32-rounded, [block][tap][c/4][n_pad][4] for fprop and the
    role-swapped layout for dgrad.  Re-packed (one kernel) whenever the dense parameter's version / storage changes."""

    def __init__(self, plan):
        nf, nd = ctypes.c_long(), ctypes.c_long()
        check(lib.hmvae_conv_packed_size(plan.handle, ctypes.byref(nf), ctypes.byref(nd)), "conv_packed_size")
        self.plan, self.nf, self.nd = plan, nf.value, nd.value
        self.wp_f = self.wp_d = None
        self.key = None
        self.event = None

    def _ensure(self, weight):
        if self.wp_f is None or self.wp_f.device != weight.device:
            self.wp_f = torch.zeros(self.nf, device=weight.device, dtype=torch.float32)   # padding stays zero forever
            self.wp_d = torch.zeros(self.nd, device=weight.device, dtype=torch.float32)
            self.key = None
            self.event = None

    def _pack(self, weight):
        check(lib.hmvae_conv_pack_weights(self.plan.handle, ptr(weight.detach()), ptr(self.wp_f), ptr(self.wp_d), stream()),
              "conv_pack_weights")
        self.key = (weight._version, weight.data_ptr())

    def prefetch(self, weight):
        """Re-packs on the CURRENT (side) stream if the weight changed and records an event that ``get`` waits for."""
        self._ensure(weight)
        if (weight._version, weight.data_ptr()) != self.key or _force_repack:
            self._pack(weight)
            self.event = torch.cuda.Event()
            self.event.record()

    def get(self, weight):
        self._ensure(weight)
        if (weight._version, weight.data_ptr()) != self.key or (_force_repack and self.event is None):
            self._pack(weight)
        if self.event is not None:                      # packed on the side stream by prefetch()
            torch.cuda.current_stream().wait_event(self.event)
            self.event = None
        return self.wp_f, self.wp_d


def prefetch_packs(items):
    """items: [(plan, weight)] in forward order.  Re-packs every changed conv weight on the side stream, so that the packing of
    layer i+1.. runs under the forward pass of layers ..i; each conv waits only for its own event."""
    if _conv_impl == IMPL_SIMT or not items:
        return
    side = _side_stream()
    side.wait_stream(torch.cuda.current_stream())         # after the optimiser update / the previous step's readers
    with torch.cuda.stream(side):
        for plan, weight in items:
            if not hasattr(plan, "packed"):
                plan.packed = PackedWeights(plan)
            plan.packed.prefetch(weight.contiguous())


def _workspace(plan, b, t_in, mode, device):
    n = int(lib.hmvae_conv_tc_workspace(plan.handle, b, t_in, mode))
    return torch.empty((n + 3) // 4, device=device, dtype=torch.float32)


def _tc_ok(plan, b, t_in, mode):
    """``plan.exact`` (set from SkeletonConv.exact): this layer always takes the fp32 CUDA-core kernels."""
    if _conv_impl == IMPL_SIMT or getattr(plan, "exact", False):
        return False
    return bool(lib.hmvae_conv_tc_supported(plan.handle, b, t_in, mode))


class _SkeletonConvFn(Function):
    """y = epilogue(conv1d(pad(prologue(x)), W (.) mask, b)) -- skeleton.py:95-105 plus the fused neighbours."""

    @staticmethod
    def forward(ctx, x, weight, bias, plan):
        x = x.contiguous()
        b, _, t_src = x.shape
        t_in = t_src * 2 if plan.upsample else t_src
        t_out = plan.t_out(t_in)
        ctot = plan.joints * plan.out_joint_stride
        shape = (b, t_out, ctot) if plan.out_channels_last else (b, ctot, t_out)
        if plan.out_joint_stride != plan.co:
            y = torch.zeros(shape, device=x.device, dtype=torch.float32)
        else:
            y = torch.empty(shape, device=x.device, dtype=torch.float32)
        w = weight.contiguous()
        tc_f, tc_d = _tc_ok(plan, b, t_in, 0), _tc_ok(plan, b, t_in, 1)
        if _conv_impl == IMPL_TC and not tc_f and not getattr(plan, "exact", False):
            raise _lib.HmvaeError("conv_fprop: the tcgen05 path does not support this geometry")
        wp_d = None
        if tc_f or tc_d:
            if not hasattr(plan, "packed"):
                plan.packed = PackedWeights(plan)
            wp_f, wp_d = plan.packed.get(w)
        if tc_f:
            ws = _workspace(plan, b, t_in, 0, x.device)
            check(lib.hmvae_conv_fprop_tc(plan.handle, ptr(x), ptr(wp_f), ptr(bias), ptr(y), b, t_in, ptr(ws), ws.numel() * 4, stream()),
                  "conv_fprop_tc")
        else:
            check(lib.hmvae_conv_fprop(plan.handle, ptr(x), ptr(w), ptr(bias), ptr(y), b, t_in, _conv_impl, stream()), "conv_fprop")
        ctx.plan, ctx.t_in, ctx.has_bias = plan, t_in, bias is not None
        ctx.bias_ref = bias.detach() if bias is not None else None      # identifies the bias gradient's arena slot
        ctx.wp_d = wp_d if tc_d else None
        ctx.save_for_backward(x, w, y if plan.lrelu else None)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, w, y = ctx.saved_tensors
        plan, t_in = ctx.plan, ctx.t_in
        gy = gy.contiguous()
        b = x.shape[0]
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            cin = plan.joints * plan.ci
            gxin = torch.empty((b, cin, t_in), device=x.device, dtype=torch.float32)
            if ctx.wp_d is not None:
                ws = _workspace(plan, b, t_in, 1, x.device)
                check(lib.hmvae_conv_dgrad_tc(plan.handle, ptr(gy), ptr(y), ptr(ctx.wp_d), ptr(gxin), b, t_in, ptr(ws), ws.numel() * 4,
                                              stream()), "conv_dgrad_tc")
            else:
                check(lib.hmvae_conv_dgrad(plan.handle, ptr(gy), ptr(y), ptr(w), ptr(gxin), b, t_in, _conv_impl, stream()), "conv_dgrad")
            if plan.has_prologue:
                gx = torch.empty_like(x)
                check(lib.hmvae_conv_prologue_bwd(plan.handle, ptr(gxin), None, ptr(gx), b, t_in, stream()), "conv_prologue_bwd")
            else:
                gx = gxin
        if ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
            # the weight gradient does not feed the rest of the backward pass: inside ``wgrad_overlap()`` it is issued on a side
            # stream so that it runs under the following layers' dgrad (most grids here are smaller than the 148 SMs)
            side = None
            if _overlap["on"]:
                side = _wgrad_stream()
                side.wait_stream(torch.cuda.current_stream())
            with (torch.cuda.stream(side) if side is not None else contextlib.nullcontext()):
                gb = grad_buffer(ctx.bias_ref) if ctx.has_bias else None
                ws = None
                if _wgrad_tc and _conv_impl != IMPL_SIMT and not getattr(plan, "exact", False) \
                        and lib.hmvae_conv_wgrad_tc_supported(plan.handle, b, t_in):
                    gw = grad_buffer(w, zero=True)            # masked blocks are never written: they must read 0
                    n = int(lib.hmvae_conv_wgrad_tc_workspace(plan.handle, b, t_in))
                    ws = torch.empty((n + 3) // 4, device=x.device, dtype=torch.float32)
                    check(lib.hmvae_conv_wgrad_tc(plan.handle, ptr(x), ptr(gy), ptr(y), ptr(gw), ptr(gb), b, t_in, 0, ptr(ws),
                                                  ws.numel() * 4, stream()), "conv_wgrad_tc")
                else:
                    gw = grad_buffer(w)
                    check(lib.hmvae_conv_wgrad(plan.handle, ptr(x), ptr(gy), ptr(y), ptr(gw), ptr(gb), b, t_in, 0, _conv_impl, stream()),
                          "conv_wgrad")
            if side is not None:
                # keep the INPUT buffers alive until the join.  gw / gb must NOT be referenced here: AccumulateGrad only steals a
                # gradient it holds the sole reference to -- otherwise it clones it on the main stream, before the side stream wrote it
                _overlap["pending"].extend([x, gy, y, ws])
        return gx, gw, gb, None


def skeleton_conv(x, weight, bias, plan):
    return _SkeletonConvFn.apply(x, weight, bias, plan)


# ------------------------------------------------------------------------------------------------ pool / unpool
class _PoolFn(Function):
    @staticmethod
    def forward(ctx, x, off, idx, c, lrelu):
        x = x.contiguous()
        b, ch, t = x.shape
        e_in, e_out = ch // c, len(off) - 1
        y = torch.empty((b, e_out * c, t), device=x.device, dtype=torch.float32)
        check(lib.hmvae_pool_fwd(ptr(x), ptr(y), b, e_in, e_out, c, t, int_array(off), int_array(idx), int(lrelu), stream()), "pool_fwd")
        ctx.meta = (off, idx, c, lrelu, e_in, e_out)
        ctx.save_for_backward(y if lrelu else None)
        return y

    @staticmethod
    def backward(ctx, gy):
        (y,) = ctx.saved_tensors
        off, idx, c, lrelu, e_in, e_out = ctx.meta
        gy = gy.contiguous()
        b, _, t = gy.shape
        gx = torch.empty((b, e_in * c, t), device=gy.device, dtype=torch.float32)
        check(lib.hmvae_pool_bwd(ptr(gy), ptr(y), ptr(gx), b, e_in, e_out, c, t, int_array(off), int_array(idx), int(lrelu), stream()),
              "pool_bwd")
        return gx, None, None, None, None


def skeleton_pool(x, pooling_list, c, lrelu=False):
    off, idx = [0], []
    for members in pooling_list:
        idx.extend(members)
        off.append(len(idx))
    return _PoolFn.apply(x, off, idx, c, lrelu)


class _UnpoolFn(Function):
    @staticmethod
    def forward(ctx, x, src, c, e_in):
        x = x.contiguous()
        b, _, t = x.shape
        e_out = len(src)
        y = torch.empty((b, e_out * c, t), device=x.device, dtype=torch.float32)
        check(lib.hmvae_unpool_fwd(ptr(x), ptr(y), b, e_in, e_out, c, t, int_array(src), stream()), "unpool_fwd")
        ctx.meta = (src, c, e_in, e_out)
        return y

    @staticmethod
    def backward(ctx, gy):
        src, c, e_in, e_out = ctx.meta
        gy = gy.contiguous()
        b, _, t = gy.shape
        gx = torch.empty((b, e_in * c, t), device=gy.device, dtype=torch.float32)
        check(lib.hmvae_unpool_bwd(ptr(gy), ptr(gx), b, e_in, e_out, c, t, int_array(src), stream()), "unpool_bwd")
        return gx, None, None, None


def skeleton_unpool(x, src, c, e_in):
    return _UnpoolFn.apply(x, src, c, e_in)


class _Upsample2Fn(Function):
    @staticmethod
    def forward(ctx, x):
        x = x.contiguous()
        t = x.shape[-1]
        y = torch.empty(x.shape[:-1] + (2 * t,), device=x.device, dtype=torch.float32)
        check(lib.hmvae_upsample2_fwd(ptr(x), ptr(y), x.numel() // t, t, stream()), "upsample2_fwd")
        return y

    @staticmethod
    def backward(ctx, gy):
        gy = gy.contiguous()
        t = gy.shape[-1] // 2
        gx = torch.empty(gy.shape[:-1] + (t,), device=gy.device, dtype=torch.float32)
        check(lib.hmvae_upsample2_bwd(ptr(gy), ptr(gx), gx.numel() // t, t, stream()), "upsample2_bwd")
        return gx


def upsample2_linear(x):
    return _Upsample2Fn.apply(x)


class _LReluFn(Function):
    @staticmethod
    def forward(ctx, x, slope):
        x = x.contiguous()
        y = torch.empty_like(x)
        check(lib.hmvae_lrelu_fwd(ptr(x), ptr(y), x.numel(), slope, stream()), "lrelu_fwd")
        ctx.slope = slope
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, gy):
        (y,) = ctx.saved_tensors
        gy = gy.contiguous()
        gx = torch.empty_like(gy)
        check(lib.hmvae_lrelu_bwd(ptr(gy), ptr(y), ptr(gx), gy.numel(), ctx.slope, stream()), "lrelu_bwd")
        return gx, None


def leaky_relu(x, slope=0.2):
    return _LReluFn.apply(x, slope)


def _transpose_raw(x):
    x = x.contiguous()
    b, c, t = x.shape
    y = torch.empty((b, t, c), device=x.device, dtype=torch.float32)
    check(lib.hmvae_transpose_ct(ptr(x), ptr(y), b, c, t, stream()), "transpose_ct")
    return y


class _TransposeFn(Function):
    @staticmethod
    def forward(ctx, x):
        return _transpose_raw(x)

    @staticmethod
    def backward(ctx, gy):
        return _transpose_raw(gy)


def transpose_ct(x):
    """[B, C, T] -> [B, T, C] (contiguous, differentiable)."""
    return _TransposeFn.apply(x)


# ------------------------------------------------------------------------------------------------ rotations / FK
class _Rot6dFn(Function):
    @staticmethod
    def forward(ctx, x6):
        x = _lib.aligned(x6)
        m = x.numel() // 6
        r = torch.empty(x.shape[:-1] + (3, 3), device=x.device, dtype=torch.float32)
        check(lib.hmvae_rot6d_fwd(ptr(x), ptr(r), m, stream()), "rot6d_fwd")
        ctx.save_for_backward(x)
        return r

    @staticmethod
    def backward(ctx, gr):
        (x,) = ctx.saved_tensors
        gr = _lib.aligned(gr)
        gx = torch.empty_like(x)
        check(lib.hmvae_rot6d_bwd(ptr(x), ptr(gr), ptr(gx), x.numel() // 6, stream()), "rot6d_bwd")
        return gx


def rot6d_to_rotmat(x6):
    return _Rot6dFn.apply(x6)


class _FKFn(Function):
    @staticmethod
    def forward(ctx, rot, offsets, positions, parents):
        rot = _lib.aligned(rot)
        rot_dim = 6 if rot.shape[-1] == 6 else 9
        n, j = rot.shape[0], rot.shape[1]
        if positions is not None:
            positions = _lib.aligned(positions)
        pos = torch.empty((n, j, 3), device=rot.device, dtype=torch.float32)
        check(lib.hmvae_fk_fwd(ptr(rot), rot_dim, ptr(offsets), ptr(positions), int_array(parents), j, n, ptr(pos), None, stream()),
              "fk_fwd")
        ctx.meta = (rot_dim, parents, n, j)
        ctx.save_for_backward(rot, offsets, positions)
        return pos

    @staticmethod
    def backward(ctx, gpos):
        rot, offsets, positions = ctx.saved_tensors
        rot_dim, parents, n, j = ctx.meta
        if ctx.needs_input_grad[2]:
            raise _lib.HmvaeError("ForwardKinematicsLayer: gradient w.r.t. the `positions` argument is not implemented")
        gpos = _lib.aligned(gpos)
        grot = torch.empty_like(rot)
        check(lib.hmvae_fk_bwd(ptr(rot), rot_dim, ptr(offsets), ptr(positions), int_array(parents), j, n, ptr(gpos), None, ptr(grot),
                               stream()), "fk_bwd")
        return grot, None, None, None


def forward_kinematics(rot, offsets, positions, parents):
    return _FKFn.apply(rot, offsets, positions, parents)


def angle_axis_to_rotation_matrix(angle_axis):
    """[N,3] -> [N,4,4] (torchgeometry semantics; no autograd -- inference / preprocessing only)."""
    aa = angle_axis.contiguous()
    if aa.dim() != 2 or aa.shape[1] != 3:
        raise ValueError("Input size must be a (*, 3) tensor. Got {}".format(tuple(aa.shape)))
    out = torch.empty((aa.shape[0], 4, 4), device=aa.device, dtype=torch.float32)
    check(lib.hmvae_aa2rot_fwd(ptr(aa), ptr(out), aa.shape[0], stream()), "aa2rot_fwd")
    return out


# ------------------------------------------------------------------------------------------------ latent heads
class _LinearFn(Function):
    @staticmethod
    def forward(ctx, x, weight, bias):
        shape = x.shape
        x2 = x.reshape(-1, shape[-1]).contiguous()
        rows, in_f, out_f = x2.shape[0], x2.shape[1], weight.shape[0]
        y = torch.empty((rows, out_f), device=x.device, dtype=torch.float32)
        w = weight.contiguous()
        check(lib.hmvae_linear_fwd(ptr(x2), ptr(w), ptr(bias), ptr(y), rows, in_f, out_f, stream()), "linear_fwd")
        ctx.save_for_backward(x2, w)
        ctx.shape, ctx.has_bias = shape, bias is not None
        ctx.bias_ref = bias.detach() if bias is not None else None
        return y.view(*shape[:-1], out_f)

    @staticmethod
    def backward(ctx, gy):
        x2, w = ctx.saved_tensors
        rows, in_f, out_f = x2.shape[0], x2.shape[1], w.shape[0]
        gy2 = gy.reshape(rows, out_f).contiguous()
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            gx = torch.empty_like(x2)
            check(lib.hmvae_linear_bwd(ptr(x2), ptr(w), ptr(gy2), ptr(gx), None, None, rows, in_f, out_f, stream()), "linear_bwd(dx)")
        if ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
            side = None
            if _overlap["on"]:
                side = _wgrad_stream()
                side.wait_stream(torch.cuda.current_stream())
            with (torch.cuda.stream(side) if side is not None else contextlib.nullcontext()):
                gw = grad_buffer(w) if ctx.needs_input_grad[1] else None
                gb = grad_buffer(ctx.bias_ref) if (ctx.has_bias and ctx.needs_input_grad[2]) else None
                check(lib.hmvae_linear_bwd(ptr(x2), ptr(w), ptr(gy2), None, ptr(gw), ptr(gb), rows, in_f, out_f, stream()), "linear_bwd(dw)")
            if side is not None:
                _overlap["pending"].extend([x2, gy2])
        return (gx.view(ctx.shape) if gx is not None else None), gw, gb


def linear(x, weight, bias=None):
    """F.linear on the library's own fp32 kernels (the latent heads of the hierarchy)."""
    return _LinearFn.apply(x, weight, bias)


# ------------------------------------------------------------------------------------------------ VAE latent
class _LatentFn(Function):
    """z = eps*exp(lv/2)+mu and the KL row-sum (seq_two_hier_sa_vae.py:419-428); dist is [rows, 2d] = (mu | lv)."""

    @staticmethod
    def forward(ctx, dist, eps, d):
        dist = dist.contiguous()
        rows = dist.numel() // (2 * d)
        z = torch.empty((rows, d), device=dist.device, dtype=torch.float32)
        kl = torch.zeros((), device=dist.device, dtype=torch.float32)
        check(lib.hmvae_latent_fwd(ptr(dist), ptr(eps), ptr(z), ptr(kl), rows, d, stream()), "latent_fwd")
        ctx.d, ctx.rows = d, rows
        ctx.save_for_backward(dist, eps)
        return z, kl

    @staticmethod
    def backward(ctx, gz, gkl):
        dist, eps = ctx.saved_tensors
        gd = torch.empty_like(dist)
        gz = gz.contiguous() if gz is not None else None
        gkl = gkl.contiguous() if gkl is not None else torch.zeros((), device=dist.device)
        check(lib.hmvae_latent_bwd(ptr(dist), ptr(eps), ptr(gz), ptr(gkl), ptr(gd), ctx.rows, ctx.d, 1.0, stream()), "latent_bwd")
        return gd, None, None


class _LatentFusedFn(Function):
    """z = eps*exp(lv/2)+mu; the KL row-sum is ACCUMULATED into ``kl_acc`` (a 1-element view of a persistent device buffer)
    and its gradient (weight ``kl_grad_scale`` = kl_w / rows, a host constant) is folded into this op's backward."""

    @staticmethod
    def forward(ctx, dist, eps, d, kl_grad_scale, kl_acc):
        dist = dist.contiguous()
        rows = dist.numel() // (2 * d)
        z = torch.empty((rows, d), device=dist.device, dtype=torch.float32)
        check(lib.hmvae_latent_fwd(ptr(dist), ptr(eps), ptr(z), kl_acc.data_ptr(), rows, d, stream()), "latent_fwd")
        ctx.d, ctx.rows, ctx.scale = d, rows, float(kl_grad_scale)
        ctx.save_for_backward(dist, eps)
        return z

    @staticmethod
    def backward(ctx, gz):
        dist, eps = ctx.saved_tensors
        gd = torch.empty_like(dist)
        check(lib.hmvae_latent_bwd(ptr(dist), ptr(eps), ptr(gz.contiguous()), None, ptr(gd), ctx.rows, ctx.d, ctx.scale, stream()),
              "latent_bwd")
        return gd, None, None, None, None


def latent_fused(dist, eps, d, kl_grad_scale, kl_acc):
    return _LatentFusedFn.apply(dist, eps, d, kl_grad_scale, kl_acc)


class _HeadsFn(Function):
    """Encoder head -> reparametrise (+ KL) -> decoder head for several hierarchy levels in ONE kernel per direction
    (``hmvae_latent_heads_fwd / _bwd``).  ``levels``: list of dicts(d, kl_scale, kl_acc, detach); tensors per level, flattened:
    (x [B, k, F], enc_w, enc_b, dec_w, dec_b, eps-or-None).  Returns the per-level decoder features [B, k, F] followed by the
    per-level distributions [B, k, 2d] (outputs for inspection; non-differentiable).
    ``detach``: the level's latent and KL are cut from the encoder (seq_two_hier_sa_vae.py:380-383): only its decoder head trains."""

    @staticmethod
    def forward(ctx, levels, *tensors):
        n = len(levels)
        ctx.set_materialize_grads(False)
        arr = (_lib.HeadLevel * n)()
        xs, feats, dists, zs, keep = [], [], [], [], []
        for l, meta in enumerate(levels):
            x, ew, eb, dw, db, eps = tensors[6 * l:6 * l + 6]
            x = x.contiguous()
            b, k, f = x.shape
            rows, d = b * k, meta["d"]
            dist = torch.empty((b, k, 2 * d), device=x.device, dtype=torch.float32)
            z = torch.empty((rows, d), device=x.device, dtype=torch.float32)
            feat = torch.empty((b, k, f), device=x.device, dtype=torch.float32)
            ew, dw = ew.contiguous(), dw.contiguous()
            h = arr[l]
            h.rows, h.features, h.d, h.kl_scale = rows, f, d, 0.0
            h.x, h.enc_w, h.enc_b, h.dec_w, h.dec_b = ptr(x), ptr(ew), ptr(eb), ptr(dw), ptr(db)
            h.eps = ptr(eps.contiguous()) if eps is not None else None
            h.dist, h.z, h.feat = ptr(dist), ptr(z), ptr(feat)
            h.kl_acc = meta["kl_acc"].data_ptr() if meta.get("kl_acc") is not None else None
            xs.append(x); feats.append(feat); dists.append(dist); zs.append(z); keep.extend([ew, dw, eps])
        check(lib.hmvae_latent_heads_fwd(arr, n, stream()), "latent_heads_fwd")
        ctx.levels = levels
        ctx.bias_refs = [(tensors[6 * l + 2], tensors[6 * l + 4]) for l in range(n)]
        ctx.save_for_backward(*xs, *dists, *zs, *[tensors[6 * l + 1] for l in range(n)], *[tensors[6 * l + 3] for l in range(n)],
                              *[t if t is not None else xs[0].new_empty(0) for t in (tensors[6 * l + 5] for l in range(n))])
        for t in dists:
            ctx.mark_non_differentiable(t)
        return tuple(feats) + tuple(dists)

    @staticmethod
    def backward(ctx, *grads):
        levels = ctx.levels
        n = len(levels)
        sv = ctx.saved_tensors
        xs, dists, zs, ews, dws, epss = (sv[i * n:(i + 1) * n] for i in range(6))
        gfeats = [g.contiguous() if g is not None else None for g in grads[:n]]
        out = [None] * (1 + 6 * n)
        live = [l for l in range(n) if gfeats[l] is not None]
        prop = [l for l in live if not levels[l].get("detach", False)]        # levels whose gradient reaches the encoder
        gdists = {}
        if prop:
            arr = (_lib.HeadLevel * len(prop))()
            for i, l in enumerate(prop):
                x, dist = xs[l], dists[l]
                b, k, f = x.shape
                d = levels[l]["d"]
                gd = torch.empty_like(dist)
                gx = torch.empty_like(x)
                h = arr[i]
                h.rows, h.features, h.d, h.kl_scale = b * k, f, d, float(levels[l]["kl_scale"])
                h.enc_w, h.dec_w, h.dist = ptr(ews[l].contiguous()), ptr(dws[l].contiguous()), ptr(dist)
                h.eps = ptr(epss[l]) if epss[l].numel() else None
                h.gfeat, h.gdist, h.gx = ptr(gfeats[l]), ptr(gd), ptr(gx)
                gdists[l] = gd
                out[1 + 6 * l] = gx
            check(lib.hmvae_latent_heads_bwd(arr, len(prop), stream()), "latent_heads_bwd")
        # weight / bias gradients of the four small linears: generic kernels, on the side stream (they feed only the optimiser)
        side = None
        if _overlap["on"]:
            side = _wgrad_stream()
            side.wait_stream(torch.cuda.current_stream())
        with (torch.cuda.stream(side) if side is not None else contextlib.nullcontext()):
            for l in live:
                x, z, gf = xs[l], zs[l], gfeats[l]
                b, k, f = x.shape
                rows, d = b * k, levels[l]["d"]
                eb, db = ctx.bias_refs[l]
                if ctx.needs_input_grad[1 + 6 * l + 3]:            # decoder head: feat = z dec_w^T + dec_b
                    gw = grad_buffer(dws[l])
                    gb = grad_buffer(db.detach()) if db is not None else None
                    check(lib.hmvae_linear_bwd(ptr(z), None, ptr(gf), None, ptr(gw), ptr(gb), rows, d, f, stream()), "linear_bwd(dw)")
                    out[1 + 6 * l + 3], out[1 + 6 * l + 4] = gw, gb
                if l in gdists and ctx.needs_input_grad[1 + 6 * l + 1]:      # encoder head: dist = x enc_w^T + enc_b
                    gw = grad_buffer(ews[l])
                    gb = grad_buffer(eb.detach()) if eb is not None else None
                    check(lib.hmvae_linear_bwd(ptr(x), None, ptr(gdists[l]), None, ptr(gw), ptr(gb), rows, f, 2 * d, stream()), "linear_bwd(dw)")
                    out[1 + 6 * l + 1], out[1 + 6 * l + 2] = gw, gb
        if side is not None:
            _overlap["pending"].extend(list(xs) + list(zs) + gfeats + list(gdists.values()))
        return tuple(out)


def latent_heads(levels, tensors):
    """See _HeadsFn.  Returns (feats per level, dists per level)."""
    res = _HeadsFn.apply(levels, *tensors)
    n = len(levels)
    return res[:n], res[n:]


def loss_finalize(acc, out, scale, w, wk):
    n = len(scale)
    arr = lambda v: (ctypes.c_float * n)(*[float(x) for x in v])
    check(lib.hmvae_loss_finalize(ptr(acc), ptr(out), arr(scale), arr(w), arr(wk), n, stream()), "loss_finalize")


def latent_sample_kl(dist, eps, d):
    """Returns (z [rows, d], kl_sum scalar).  kl mean = kl_sum / rows."""
    return _LatentFn.apply(dist, eps, d)


# ------------------------------------------------------------------------------------------------ MSE
class _MseSumFn(Function):
    @staticmethod
    def forward(ctx, a, b):
        a, b = a.contiguous(), b.contiguous()
        out = torch.zeros((), device=a.device, dtype=torch.float32)
        check(lib.hmvae_mse_fwd(ptr(a), ptr(b), ptr(out), a.numel(), stream()), "mse_fwd")
        ctx.save_for_backward(a, b)
        return out

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        da = torch.empty_like(a)
        check(lib.hmvae_mse_bwd(ptr(a), ptr(b), ptr(da), a.numel(), 2.0, stream()), "mse_bwd")
        return da * g, None


def l2_criterion(pred, gt):
    """mean((pred-gt)^2) -- seq_two_hier_sa_vae.py:430-434 (gt gets no gradient)."""
    assert pred.size() == gt.size()
    return _MseSumFn.apply(pred, gt) / pred.numel()


# ------------------------------------------------------------------------------------------------ fused losses / optimiser
def recon_fwdbwd(x6_pred, ncw, gt_6d, gt_rotmat, offsets, parents, w6d, wrot, wpos, losses, want_grad=True, mask=None,
                 rot_out=None, pos_out=None):
    """Fused GT-FK + rot6d + FK + 3 MSE (+ gradient w.r.t. x6_pred).  `losses` is a zeroed float32[>=3] device tensor
    that receives the three squared-error SUMS.  Returns dx6 (same layout as x6_pred) or None.
    ``mask`` [B, T, joints]: l2_masked_criterion weights (seq_two_hier_sa_vae.py:717-735); ``rot_out`` [B,T,J,3,3] / ``pos_out``
    [B,T,J,3]: optional outputs of the predicted rotation matrices / FK positions."""
    x6_pred = x6_pred.contiguous()
    if ncw:
        b, c, t = x6_pred.shape
    else:
        b, t, c = x6_pred.shape
    j = c // 6
    nf = float(b * t)
    dx6 = torch.empty_like(x6_pred) if want_grad else None
    if mask is not None:
        mask = mask.to(dtype=torch.float32).contiguous()
        if mask.numel() != b * t * j:
            raise ValueError("mask must be [B, T, joints]")
    check(lib.hmvae_recon_masked_fwdbwd(ptr(x6_pred), int(ncw), ptr(_lib.aligned(gt_6d)), ptr(_lib.aligned(gt_rotmat)), ptr(mask),
                                        ptr(offsets), int_array(parents), j, b, t, 2.0 * w6d / (nf * 6 * j),
                                        2.0 * wrot / (nf * 9 * j), 2.0 * wpos / (nf * 3 * j), ptr(losses), ptr(dx6), ptr(pos_out),
                                        None, ptr(rot_out), stream()), "recon_fwdbwd")
    return dx6


def l2_reg_fwdbwd(params, refs, weight, loss, grads=None, accumulate=None):
    """loss[0] += sum_i mean((p_i - ref_i)^2); grads[i] (+)= weight * 2 (p_i - ref_i) / numel_i  (one launch).
    ``grads[i]`` may be None (value only); ``accumulate[i]``: add to an existing gradient instead of storing."""
    n = len(params)
    arr = (_lib.RegTensor * n)()
    for i, (p, r) in enumerate(zip(params, refs)):
        g = grads[i] if grads is not None else None
        arr[i] = _lib.RegTensor(p.data_ptr(), r.data_ptr(), g.data_ptr() if g is not None else None, p.numel(),
                                int(bool(accumulate[i])) if accumulate is not None else 0)
    check(lib.hmvae_l2_reg_fwdbwd(arr, n, float(weight), ptr(loss), stream()), "l2_reg_fwdbwd")


def traj_fwdbwd(root_v_pred, root_v_gt, mean3, std3, joints, w_v, w_trans, losses, want_grad=True):
    root_v_pred, root_v_gt = root_v_pred.contiguous(), root_v_gt.contiguous()
    b, t, _ = root_v_pred.shape
    d = torch.empty_like(root_v_pred) if want_grad else None
    m = (ctypes.c_float * 3)(*[float(v) for v in mean3])
    s = (ctypes.c_float * 3)(*[float(v) for v in std3])
    check(lib.hmvae_traj_fwdbwd(ptr(root_v_pred), ptr(root_v_gt), m, s, b, t, joints, 2.0 * w_v / (b * t * 3),
                                2.0 * w_trans / (t * b * joints * 3), ptr(losses), ptr(d), stream()), "traj_fwdbwd")
    return d


class FusedAdam:
    """torch.optim.Adam(lr, betas, eps, weight_decay) semantics in one multi-tensor kernel (trainer_motion_vae.py:29-31), with
    torch.optim.lr_scheduler.StepLR folded in (``set_schedule``).  The step counter and the schedule position live on the device
    (``hmvae_opt_clock_tick``): a replayed CUDA graph of the step keeps its own time."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        self.params = [p for p in params]
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.step_count = 0            # host mirror of clock[0] (advance() is called once per step)
        self.sched_iters = 0           # host mirror of clock[1]
        self.gamma, self.step_size = 1.0, 0
        self.exp_avg = [torch.zeros_like(p) for p in self.params]
        self.exp_avg_sq = [torch.zeros_like(p) for p in self.params]
        self.param_groups = [dict(lr=lr)]
        self._live = set()
        dev = self.params[0].device
        self._clock = torch.zeros(2, dtype=torch.int32, device=dev)
        self._dyn_dev = torch.zeros(2, dtype=torch.float32, device=dev)

    def set_schedule(self, gamma, step_size):
        """StepLR(step_size, gamma) evaluated on the device; step_size <= 0 = constant learning rate."""
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

    def zero_grad(self, set_to_none=True):
        for p in self.params:
            if set_to_none:
                p.grad = None
            elif p.grad is not None:
                p.grad.zero_()

    def _pack(self):
        live = [(i, p, m, v) for i, (p, m, v) in enumerate(zip(self.params, self.exp_avg, self.exp_avg_sq)) if p.grad is not None]
        arr = (_lib.AdamTensor * len(live))()
        for k, (i, p, m, v) in enumerate(live):
            g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
            arr[k] = _lib.AdamTensor(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel())
            self._live.add(i)
        return arr, len(live)

    def _bump_versions(self):
        # the kernels update parameters through raw pointers: tell autograd / the packed-weight caches
        for p in self.params:
            if p.grad is not None:
                torch.autograd.graph.increment_version(p)

    def step(self, grad_scale=1.0, lr=None, step=None):
        """Eager step with host-side scalars (tests / simple loops)."""
        self.step_count = self.step_count + 1 if step is None else step
        self.sched_iters += 1
        arr, n = self._pack()
        self._bump_versions()
        check(lib.hmvae_adam_step(arr, n, self.param_groups[0]["lr"] if lr is None else lr, self.betas[0], self.betas[1],
                                  self.eps, self.weight_decay, self.step_count, grad_scale, stream()), "adam_step")
        self._clock.copy_(torch.tensor([self.step_count, self.sched_iters], dtype=torch.int32))

    # ---- CUDA-graph friendly variant: the step-dependent scalars are produced on the device
    def advance(self, lr=None):
        """Host side of one step: only the mirrors move (the device clock ticks inside the step)."""
        self.param_groups[0]["lr"] = self.current_lr()
        self.step_count += 1
        self.sched_iters += 1

    def tick(self):
        """Device side (capturable): advance the clock and refresh {lr/(1-b1^t), 1/sqrt(1-b2^t)}."""
        check(lib.hmvae_opt_clock_tick(self._clock.data_ptr(), self.lr, self.gamma, self.step_size, self.betas[0], self.betas[1],
                                       ptr(self._dyn_dev), stream()), "opt_clock_tick")

    def step_dyn(self, grad_scale=1.0):
        """Device side (capturable): clock tick + the multi-tensor kernel.  Call advance() first."""
        self.tick()
        arr, n = self._pack()
        self._bump_versions()
        check(lib.hmvae_adam_step_dyn(arr, n, ptr(self._dyn_dev), self.betas[0], self.betas[1], self.eps, self.weight_decay, grad_scale,
                                      stream()), "adam_step_dyn")

    def state_dict(self):
        """torch.optim.Adam.state_dict() layout (the reference's optimizer.pt, trainer_motion_vae.py:112-113)."""
        from .optim_state import to_torch_adam

        return to_torch_adam(self.step_count, self.current_lr(), self.betas, self.eps, self.weight_decay, self.exp_avg,
                             self.exp_avg_sq, live=self._live, initial_lr=self.lr)

    def load_state_dict(self, sd):
        from .optim_state import from_torch_adam

        step, lr, m, v, live = from_torch_adam(sd, len(self.params))
        for dst, src in zip(self.exp_avg, m):
            dst.zero_() if src is None else dst.copy_(src)
        for dst, src in zip(self.exp_avg_sq, v):
            dst.zero_() if src is None else dst.copy_(src)
        self._live = set(live)
        self.set_clock(step, self.sched_iters)
