"""Stack-level execution of the hierarchical encoder / decoder (seq_two_hier_sa_vae.py:120-130, 142-167, 233-294).

The per-layer path (ops.skeleton_conv / skeleton_pool) runs, for every conv, a staging pass, the tcgen05 kernel, a finish pass
and -- between layers -- a pool / unpool-upsample-adjoint kernel: 3-4 dependent launches per layer boundary, which at B=32 cost
more than the tensor-core kernels themselves (tools/tc_phases.py).  Here a whole stack is ONE autograd node and neighbouring
convs are joined by ``hmvae_conv_link``:

    forward :  stage(conv 0) -> run 0 -> link -> run 1 -> link -> ... -> run n-1 -> link (no consumer)
    backward:  stage(dgrad n-1) -> run -> link -> run n-2 -> ... -> link (no consumer);  weight gradients on the side stream

Every link writes the boundary tensor (activated / pooled layer output, or its gradient) exactly once -- the backward pass and
the weight-gradient kernels need it anyway -- and scatters it into the next conv's staged tiles.  Falls back to the per-layer
path (returns None from ``*_forward``) when a geometry is not supported by the tensor-core kernels or the link.
"""
import contextlib
import ctypes
import os
import weakref

import torch
from torch.autograd import Function

from . import _lib, ops
from ._lib import check, int_array, lib, ptr, stream

_enabled = True
_specs = weakref.WeakKeyDictionary()       # module -> {(batch, length, device): spec}; kept out of the module's __dict__ (deepcopy)
_last_plans = weakref.WeakKeyDictionary()  # decoder -> plans of its last stack run (for ops.prefetch_packs)


def set_enabled(flag):
    """False: Encoder / Decoder use the per-layer path (A/B comparisons, parity tests)."""
    global _enabled
    _enabled = bool(flag)


def _bufs(plan, mode, b, t):
    """Persistent staging buffer (zero-initialised once: the link kernel never writes padding) and accumulator-dump buffer of
    one (plan, mode, batch, length)."""
    cache = plan.__dict__.setdefault("_stack_bufs", {})
    key = (mode, b, t, torch.cuda.current_device())
    got = cache.get(key)
    if got is None:
        sb, db = ctypes.c_long(), ctypes.c_long()
        if not lib.hmvae_conv_tc_sizes(plan.handle, b, t, mode, ctypes.byref(sb), ctypes.byref(db)):
            return None
        dev = torch.device("cuda", torch.cuda.current_device())
        got = (torch.zeros((sb.value + 3) // 4, device=dev, dtype=torch.float32), torch.empty((db.value + 3) // 4, device=dev, dtype=torch.float32))
        cache[key] = got
    return got


def _csr(lists):
    off, idx = [0], []
    for members in lists:
        idx.extend(int(v) for v in members)
        off.append(len(idx))
    return off, idx


def _link_desc(kind, b, prod, prod_t, cons=None, cons_t=0, act=False, pool=None, dump=None, bias=None, aux=None, add=None, sact=None,
               yact_c=None, s_out=None, stage_ws=None):
    d = _lib.ConvLinkDesc()
    d.kind, d.batch, d.prod, d.prod_t = kind, b, prod.handle, prod_t
    d.cons, d.cons_t = (cons.handle if cons is not None else None), cons_t
    d.act = int(bool(act))
    keep = []
    if pool is not None:
        off, idx = _csr(pool)
        keep = [int_array(off), int_array(idx)]
        d.pool_joints, d.pool_off, d.pool_idx = len(pool), keep[0], keep[1]
    for name, t in (("dump", dump), ("bias", bias), ("aux", aux), ("add", add), ("sact", sact), ("yact_c", yact_c), ("s_out", s_out),
                    ("stage_ws", stage_ws)):
        setattr(d, name, ptr(t) if t is not None else None)
    d._keep = keep
    return d


def _link(**kw):
    d = _link_desc(**kw)
    check(lib.hmvae_conv_link(ctypes.byref(d), stream()), "conv_link")


def _link_ok(**kw):
    return bool(lib.hmvae_conv_link_supported(ctypes.byref(_link_desc(**kw))))


def _packed(plan, w):
    if not hasattr(plan, "packed"):
        plan.packed = ops.PackedWeights(plan)
    return plan.packed.get(w)


def _is_identity(pooling_list):
    return all(len(p) == 1 and p[0] == k for k, p in enumerate(pooling_list))


class _Geometry:
    """Shapes of one stack for a batch size: per conv the virtual input length, per boundary the tensor shape."""

    def __init__(self, plans, t_src0):
        self.t_in, self.t_out = [], []
        t = t_src0
        for p in plans:
            t_in = 2 * t if p.upsample else t
            self.t_in.append(t_in)
            t = p.t_out(t_in)
            self.t_out.append(t)


_prestage = os.environ.get("HMVAE_WGRAD_PRESTAGE", "1") != "0"
# under no_grad the boundary tensors between two convs that nobody reads are not written (the link kernel only stages the consumer)
_skip_bounds = os.environ.get("HMVAE_STACK_SKIP_BOUNDS", "1") != "0"


def _prestage_x(plan, x, b, t_in, want, after=None):
    """The x operand of conv ``plan``'s tensor-core weight gradient depends on forward data only: its TF32 tiles are staged NOW,
    on a side stream beside the forward chain (which leaves most of the GPU idle), instead of inside the crowded backward pass.
    ``after``: event recorded when ``x`` was complete (default: the current stream's position).  Returns the workspace the
    weight-gradient call receives (with x = NULL), or None."""
    if not (want and _prestage and ops._wgrad_tc and ops._overlap["allowed"] and lib.hmvae_conv_wgrad_tc_supported(plan.handle, b, t_in)):
        return None
    ps = ops._prestage_stream()
    if after is not None:
        ps.wait_event(after)
    else:
        ps.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(ps):
        nbytes = int(lib.hmvae_conv_wgrad_tc_workspace(plan.handle, b, t_in))
        ws = torch.empty((nbytes + 3) // 4, device=x.device, dtype=torch.float32)
        check(lib.hmvae_conv_wgrad_tc_stage_x(plan.handle, ptr(x), b, t_in, ptr(ws), ws.numel() * 4, stream()), "conv_wgrad_tc_stage_x")
    return ws


def _mark():
    ev = torch.cuda.Event()
    ev.record()
    return ev


def _wgrad(plan, x, gy, y, w, bias, b, t_in, want_w, want_b, inline=False, pre=None):
    """Weight / bias gradient of one conv on the side stream (inside ops.wgrad_overlap) -- same kernels as the per-layer path.
    ``inline``: on the caller's stream instead (the LAST weight gradient of a backward pass: nothing is left to overlap with, and
    round-robin would queue it behind an earlier layer's kernels on a side stream)."""
    if not (want_w or want_b):
        return None, None
    side = None
    if ops._overlap["on"] and not inline:
        side = ops._wgrad_stream()
        side.wait_stream(torch.cuda.current_stream())
    if pre is not None:                                   # x tiles staged during the forward pass (_prestage_x)
        (side if side is not None else torch.cuda.current_stream()).wait_stream(ops._prestage_stream())
    with (torch.cuda.stream(side) if side is not None else contextlib.nullcontext()):
        gb = ops.grad_buffer(bias.detach()) if (bias is not None and want_b) else None
        ws = None
        if pre is not None or (ops._wgrad_tc and lib.hmvae_conv_wgrad_tc_supported(plan.handle, b, t_in)):
            gw = ops.grad_buffer(w, zero=True)
            if pre is not None:
                ws = pre
            else:
                n = int(lib.hmvae_conv_wgrad_tc_workspace(plan.handle, b, t_in))
                ws = torch.empty((n + 3) // 4, device=x.device, dtype=torch.float32)
            check(lib.hmvae_conv_wgrad_tc(plan.handle, None if pre is not None else ptr(x), ptr(gy), ptr(y), ptr(gw), ptr(gb), b, t_in, 0,
                                          ptr(ws), ws.numel() * 4, stream()), "conv_wgrad_tc")
        else:
            gw = ops.grad_buffer(w)
            check(lib.hmvae_conv_wgrad(plan.handle, ptr(x), ptr(gy), ptr(y), ptr(gw), ptr(gb), b, t_in, 0, 0, stream()), "conv_wgrad")
    if side is not None:
        ops._overlap["pending"].extend([x, gy, y, ws])
    return gw, gb


# ======================================================================================================== decoder
class _DecoderStackFn(Function):
    @staticmethod
    def forward(ctx, spec, feat0, feat_last, *params):
        plans, geo, b = spec["plans"], spec["geo"], feat0.shape[0]
        n = len(plans)
        ctx.set_materialize_grads(False)
        ws, bs = params[:n], params[n:]                      # weights, biases (None where the conv has none)
        feat0 = feat0.contiguous()
        concat = feat_last is not None
        if concat:
            feat_last = feat_last.contiguous()
        dev = feat0.device
        # ctx.needs_input_grad mirrors requires_grad of the inputs whatever the grad mode: the caller's grad mode rides in the spec
        spec_grad = spec["grad"] and any(ctx.needs_input_grad)
        bounds = [feat0]                                     # bounds[i] = source tensor of conv i
        st, dump = _bufs(plans[0], 0, b, geo.t_in[0])
        check(lib.hmvae_conv_tc_stage(plans[0].handle, 0, ptr(feat0), None, b, geo.t_in[0], ptr(st), stream()), "conv_tc_stage")
        packs = [None] * n                                   # each conv waits for ITS packed weights only (packing runs on the side stream)
        packs[0] = _packed(plans[0], ws[0])
        check(lib.hmvae_conv_tc_run(plans[0].handle, 0, ptr(packs[0][0]), b, geo.t_in[0], ptr(st), ptr(dump), stream()), "conv_tc_run")
        pre = [None] * n                                     # weight-gradient x tiles, staged beside the forward chain
        if spec_grad:
            pre[0] = _prestage_x(plans[0], feat0, b, geo.t_in[0], ctx.needs_input_grad[3])
        for i in range(1, n + 1):
            prod = plans[i - 1]
            cons = plans[i] if i < n else None
            chans = prod.joints * prod.out_joint_stride
            # interior boundary tensors are only needed again by the backward pass: not written under no_grad
            keep = cons is None or spec_grad or not _skip_bounds
            s_i = torch.empty((b, chans, geo.t_out[i - 1]), device=dev, dtype=torch.float32) if keep else None
            aux = feat_last if (concat and i == n - 1) else None
            if cons is not None:
                st_c, dump_c = _bufs(cons, 0, b, geo.t_in[i])
            _link(kind=0, b=b, prod=prod, prod_t=geo.t_in[i - 1], cons=cons, cons_t=geo.t_in[i] if cons is not None else 0,
                  act=prod.lrelu, dump=dump, bias=bs[i - 1], aux=aux, s_out=s_i, stage_ws=st_c if cons is not None else None)
            bounds.append(s_i)
            if cons is not None:
                ready = _mark() if (spec_grad and _prestage and ctx.needs_input_grad[3 + i]) else None
                packs[i] = _packed(cons, ws[i])
                check(lib.hmvae_conv_tc_run(cons.handle, 0, ptr(packs[i][0]), b, geo.t_in[i], ptr(st_c), ptr(dump_c), stream()), "conv_tc_run")
                dump = dump_c
                if ready is not None:                        # issued after the chain's kernel: issue order = dispatch order
                    pre[i] = _prestage_x(cons, s_i, b, geo.t_in[i], True, after=ready)
        if spec_grad:
            ctx.wg_pre = pre
            ctx.spec, ctx.b, ctx.concat = spec, b, concat
            ctx.packs_d = [p[1] for p in packs]
            ctx.has_bias = [x is not None for x in bs]
            ctx.save_for_backward(*bounds, *ws, *[x for x in bs if x is not None])
        return bounds[n]

    @staticmethod
    def backward(ctx, gout):
        spec, b = ctx.spec, ctx.b
        plans, geo = spec["plans"], spec["geo"]
        n = len(plans)
        if gout is None:
            if any(w is not None for w in ctx.wg_pre):
                torch.cuda.current_stream().wait_stream(ops._prestage_stream())
            return (None,) * (3 + 2 * n)
        saved = ctx.saved_tensors
        bounds, ws = saved[:n + 1], saved[n + 1:2 * n + 1]
        bias_iter = iter(saved[2 * n + 1:])
        bs = [next(bias_iter) if hb else None for hb in ctx.has_bias]
        gout = gout.contiguous()
        dev = gout.device
        need_w = [ctx.needs_input_grad[3 + i] for i in range(n)]
        need_b = [bs[i] is not None and ctx.needs_input_grad[3 + n + i] for i in range(n)]
        need_x = ctx.needs_input_grad[1] or (ctx.concat and ctx.needs_input_grad[2])
        gws, gbs = [None] * n, [None] * n
        gbounds = [None] * (n + 1)
        gbounds[n] = gout
        # the last conv's weight gradient needs nothing from the dgrad chain.  Issued AFTER the chain's first two kernels: nodes of a
        # replayed graph that become ready together are dispatched in issue order, and the staging pass of this weight gradient
        # fills every SM's thread slots for ~30 us (tools/timeline.py: the first dgrad kernel started 10 us late behind it)
        chain = need_x or any(need_w[:-1]) or any(need_b[:-1])
        pre = ctx.wg_pre
        first = lambda: _wgrad(plans[n - 1], bounds[n - 1], gout, bounds[n] if plans[n - 1].lrelu else None, ws[n - 1], bs[n - 1],
                               b, geo.t_in[n - 1], need_w[n - 1], need_b[n - 1], pre=pre[n - 1])
        if not (chain and _wgrad_after_dgrad):
            gws[n - 1], gbs[n - 1] = first()
        if chain:
            st, dump = _bufs(plans[n - 1], 1, b, geo.t_in[n - 1])
            check(lib.hmvae_conv_tc_stage(plans[n - 1].handle, 1, ptr(gout), ptr(bounds[n]) if plans[n - 1].lrelu else None, b,
                                          geo.t_in[n - 1], ptr(st), stream()), "conv_tc_stage")
            check(lib.hmvae_conv_tc_run(plans[n - 1].handle, 1, ptr(ctx.packs_d[n - 1]), b, geo.t_in[n - 1], ptr(st), ptr(dump), stream()),
                  "conv_tc_run")
            if _wgrad_after_dgrad:
                gws[n - 1], gbs[n - 1] = first()
            for i in range(n - 1, -1, -1):
                prod = plans[i]
                cons = plans[i - 1] if i > 0 else None
                g_i = torch.empty_like(bounds[i])
                if cons is not None:
                    st_c, dump_c = _bufs(cons, 1, b, geo.t_in[i - 1])
                _link(kind=1, b=b, prod=prod, prod_t=geo.t_in[i], cons=cons, cons_t=geo.t_in[i - 1] if cons is not None else 0,
                      dump=dump, yact_c=bounds[i] if (cons is not None and cons.lrelu) else None, s_out=g_i,
                      stage_ws=st_c if cons is not None else None)
                gbounds[i] = g_i
                if cons is not None:
                    check(lib.hmvae_conv_tc_run(cons.handle, 1, ptr(ctx.packs_d[i - 1]), b, geo.t_in[i - 1], ptr(st_c), ptr(dump_c), stream()),
                          "conv_tc_run")
                    dump = dump_c
                    gws[i - 1], gbs[i - 1] = _wgrad(cons, bounds[i - 1], g_i, bounds[i] if cons.lrelu else None, ws[i - 1], bs[i - 1], b,
                                                    geo.t_in[i - 1], need_w[i - 1], need_b[i - 1], pre=pre[i - 1])
        gfeat0 = gbounds[0]
        gfeat_last = None
        if ctx.concat and gbounds[n - 1] is not None:
            e = plans[n - 2].joints
            g = gbounds[n - 1]
            c2 = g.shape[1] // e
            cp = plans[n - 2].co
            gfeat_last = g.view(b, e, c2, g.shape[2])[:, :, cp:, :].reshape(b, e * (c2 - cp), g.shape[2])
        return (None, gfeat0, gfeat_last) + tuple(gws) + tuple(gbs)


def decoder_spec(dec, b, t_src0):
    """Plans + geometry of the decoder stack, or None if the stack path cannot run it."""
    hp = dec.hp
    n = hp['num_layers']
    if hp['extra_conv'] or n < 2 or any(getattr(c, "exact", False) for c in dec.convs):
        return None
    key = ("dec", b, t_src0, torch.cuda.current_device())
    cache = _specs.setdefault(dec, {})
    if key in cache:
        return cache[key]
    plans = []
    for i, conv in enumerate(dec.convs):
        kw = dec._fused_kwargs(i)
        if i == n - 2:
            kw["out_joint_stride"] = 2 * conv.out_channels_per_joint       # its output is the first half of the per-edge concat
        plans.append(conv.plan(**kw))
    geo = _Geometry(plans, t_src0)
    ok = True
    for i in range(n):
        ok = ok and _bufs(plans[i], 0, b, geo.t_in[i]) is not None and _bufs(plans[i], 1, b, geo.t_in[i]) is not None
    if ok:
        dummy_aux = torch.empty(1, device="cuda")
        for i in range(1, n + 1):
            cons = plans[i] if i < n else None
            ok = ok and _link_ok(kind=0, b=b, prod=plans[i - 1], prod_t=geo.t_in[i - 1], cons=cons, cons_t=geo.t_in[i] if cons else 0,
                                 act=plans[i - 1].lrelu, aux=dummy_aux if i == n - 1 else None)
            consb = plans[i - 2] if i >= 2 else None
            ok = ok and _link_ok(kind=1, b=b, prod=plans[i - 1], prod_t=geo.t_in[i - 1], cons=consb, cons_t=geo.t_in[i - 2] if consb else 0,
                                 yact_c=dummy_aux if (consb is not None and consb.lrelu) else None)
    spec = dict(plans=plans, geo=geo) if ok else None
    cache[key] = spec
    return spec


def decoder_forward(dec, feat0, feat_last):
    """Decoder.forward's conv stack on (hier_feats[0], hier_feats[n-1]) -> bs X (24*6) X T, or None (caller falls back)."""
    if not _enabled or ops._conv_impl == ops.IMPL_SIMT:
        return None
    b, t0 = feat0.shape[0], feat0.shape[2]
    spec = decoder_spec(dec, b, t0)
    if spec is None:
        return None
    _last_plans[dec] = spec["plans"]
    ws = [c.weight for c in dec.convs]
    bs = [c.bias for c in dec.convs]
    return _DecoderStackFn.apply(dict(spec, grad=torch.is_grad_enabled()), feat0, feat_last, *ws, *bs)


# ======================================================================================================== encoder
# encoder module -> optional callback(layer, n_layers), called from inside that encoder stack's backward right after the weight /
# bias gradient kernels of ``layer`` have been issued (they write into the registered gradient buffers): lets the optimiser start
# on that bucket while the shallower levels' backward still runs
wgrad_issued_hooks = weakref.WeakKeyDictionary()
_last_inline = os.environ.get("HMVAE_WGRAD_LAST_INLINE", "1") != "0"
_wgrad_after_dgrad = os.environ.get("HMVAE_WGRAD_AFTER_DGRAD", "1") != "0"


class _EncoderStackFn(Function):
    @staticmethod
    def forward(ctx, spec, x, *params):
        plans, geo, pools, b = spec["plans"], spec["geo"], spec["pools"], x.shape[0]
        n = len(plans)
        ctx.set_materialize_grads(False)      # levels without a latent head get None, not a zero tensor (3 fill kernels + 3 dead adds)
        ws, bs = params[:n], params[n:]
        x = x.contiguous()
        dev = x.device
        spec_grad = spec["grad"] and any(ctx.needs_input_grad)
        bounds = [x]
        st, dump = _bufs(plans[0], 0, b, geo.t_in[0])
        check(lib.hmvae_conv_tc_stage(plans[0].handle, 0, ptr(x), None, b, geo.t_in[0], ptr(st), stream()), "conv_tc_stage")
        packs = [None] * n                                   # each conv waits for ITS packed weights only (packing runs on the side stream)
        packs[0] = _packed(plans[0], ws[0])
        check(lib.hmvae_conv_tc_run(plans[0].handle, 0, ptr(packs[0][0]), b, geo.t_in[0], ptr(st), ptr(dump), stream()), "conv_tc_run")
        pre = [None] * n                                     # weight-gradient x tiles, staged beside the forward chain
        if spec_grad:
            pre[0] = _prestage_x(plans[0], x, b, geo.t_in[0], ctx.needs_input_grad[2])
        for i in range(1, n + 1):
            prod = plans[i - 1]
            cons = plans[i] if i < n else None
            pool = pools[i - 1]
            e_out = len(pool) if pool is not None else prod.joints
            keep = cons is None or spec_grad or (i - 1) in spec["needed"] or not _skip_bounds
            s_i = torch.empty((b, e_out * prod.co, geo.t_out[i - 1]), device=dev, dtype=torch.float32) if keep else None
            if cons is not None:
                st_c, dump_c = _bufs(cons, 0, b, geo.t_in[i])
            _link(kind=0, b=b, prod=prod, prod_t=geo.t_in[i - 1], cons=cons, cons_t=geo.t_in[i] if cons is not None else 0, act=True,
                  pool=pool, dump=dump, bias=bs[i - 1], s_out=s_i, stage_ws=st_c if cons is not None else None)
            bounds.append(s_i)
            if cons is not None:
                ready = _mark() if (spec_grad and _prestage and ctx.needs_input_grad[2 + i]) else None
                packs[i] = _packed(cons, ws[i])
                check(lib.hmvae_conv_tc_run(cons.handle, 0, ptr(packs[i][0]), b, geo.t_in[i], ptr(st_c), ptr(dump_c), stream()), "conv_tc_run")
                dump = dump_c
                if ready is not None:
                    pre[i] = _prestage_x(cons, s_i, b, geo.t_in[i], True, after=ready)
        if spec_grad:
            ctx.wg_pre = pre
            ctx.spec, ctx.b = spec, b
            ctx.packs_d = [p[1] for p in packs]
            ctx.has_bias = [v is not None for v in bs]
            ctx.save_for_backward(*bounds, *ws, *[v for v in bs if v is not None])
        return tuple(t if t is not None else x.new_empty(0) for t in bounds[1:])

    @staticmethod
    def backward(ctx, *gs):
        spec, b = ctx.spec, ctx.b
        plans, geo, pools = spec["plans"], spec["geo"], spec["pools"]
        n = len(plans)
        saved = ctx.saved_tensors
        bounds, ws = saved[:n + 1], saved[n + 1:2 * n + 1]
        bias_iter = iter(saved[2 * n + 1:])
        bs = [next(bias_iter) if hb else None for hb in ctx.has_bias]
        need_w = [ctx.needs_input_grad[2 + i] for i in range(n)]
        need_b = [bs[i] is not None and ctx.needs_input_grad[2 + n + i] for i in range(n)]
        gs = [g.contiguous() if g is not None else None for g in gs]        # gs[i-1] = gradient of bounds[i] from the latent heads
        gws, gbs = [None] * n, [None] * n
        if ctx.needs_input_grad[1]:
            raise _lib.HmvaeError("encoder stack: gradient w.r.t. the network input is not implemented on the stack path")
        live = [i for i in range(n) if gs[i] is not None]
        top = live[-1] if live else -1
        if any(w is not None for w in ctx.wg_pre[top + 1:]):      # pre-staged tiles nobody will consume: re-join their stream
            torch.cuda.current_stream().wait_stream(ops._prestage_stream())
        if not live:
            return (None, None) + tuple(gws) + tuple(gbs)
        # Notation: R_i = raw output of conv i, bounds[i+1] = LeakyReLU(pool_i(R_i)).  The chain starts at the deepest conv whose
        # activated output received a gradient (convs above it get no weight gradient, like torch leaves .grad = None there).
        top = live[-1]
        dev = gs[top].device
        if pools[top] is None:
            gy, yact = gs[top], bounds[top + 1]           # LeakyReLU is the conv's own epilogue: the staging pass applies LeakyReLU'
        else:
            off, idx = _csr(pools[top])
            pt = plans[top]
            gy = torch.empty((b, pt.joints * pt.co, geo.t_out[top]), device=dev, dtype=torch.float32)
            check(lib.hmvae_pool_bwd(ptr(gs[top]), ptr(bounds[top + 1]), ptr(gy), b, pt.joints, len(pools[top]), pt.co, geo.t_out[top],
                                     int_array(off), int_array(idx), 1, stream()), "pool_bwd")
            yact = None
        staged = False
        for i in range(top, -1, -1):
            plan = plans[i]

            def wgrad_i(i=i, plan=plan, gy=gy, yact=yact):
                gws[i], gbs[i] = _wgrad(plan, bounds[i], gy, yact, ws[i], bs[i], b, geo.t_in[i], need_w[i], need_b[i],
                                        inline=(i == 0 and _last_inline), pre=ctx.wg_pre[i])
                if spec.get("hook") is not None:
                    spec["hook"](i, n)                     # e.g. the optimiser's bucket for this level (Trainer, split step)
            if i == 0 or not _wgrad_after_dgrad:
                wgrad_i()
            if i == 0:
                break                                      # the network input has no gradient: conv 0 needs no dgrad
            st, dump = _bufs(plan, 1, b, geo.t_in[i])
            if not staged:
                check(lib.hmvae_conv_tc_stage(plan.handle, 1, ptr(gy), ptr(yact), b, geo.t_in[i], ptr(st), stream()), "conv_tc_stage")
            check(lib.hmvae_conv_tc_run(plan.handle, 1, ptr(ctx.packs_d[i]), b, geo.t_in[i], ptr(st), ptr(dump), stream()), "conv_tc_run")
            if _wgrad_after_dgrad:
                wgrad_i()                                  # after the chain's kernel of this level (issue order = dispatch order)
            prev = plans[i - 1]
            cons = prev if i - 1 >= 1 else None
            gy_prev = torch.empty((b, prev.joints * prev.co, geo.t_out[i - 1]), device=dev, dtype=torch.float32)
            st_c = _bufs(cons, 1, b, geo.t_in[i - 1])[0] if cons is not None else None
            # gradient of conv i-1's raw output: pool^T(LeakyReLU'(bounds[i]) * (dgrad of conv i + head gradient)), staged for dgrad i-1
            _link(kind=2, b=b, prod=plan, prod_t=geo.t_in[i], cons=cons, cons_t=geo.t_in[i - 1] if cons is not None else 0, act=True,
                  pool=pools[i - 1], dump=dump, add=gs[i - 1], sact=bounds[i], s_out=gy_prev, stage_ws=st_c)
            staged = True
            gy, yact = gy_prev, None
        return (None, None) + tuple(gws) + tuple(gbs)


def encoder_spec(enc, b, t0):
    if enc.args['extra_conv'] or any(getattr(c, "exact", False) for c in enc.convs):
        return None
    key = ("enc", b, t0, torch.cuda.current_device())
    cache = _specs.setdefault(enc, {})
    if key in cache:
        return cache[key]
    n = len(enc.convs)
    pools = [None if _is_identity(enc.pools[i].pooling_list) else [list(p) for p in enc.pools[i].pooling_list] for i in range(n)]
    # interior identity pools would put LeakyReLU into the conv epilogue AND need the kind-2 link: keep it simple, last level only
    if any(pools[i] is None for i in range(n - 1)):
        cache[key] = None
        return None
    plans = [c.plan(lrelu=True) if pools[i] is None else c.plan() for i, c in enumerate(enc.convs)]
    geo = _Geometry(plans, t0)
    ok = True
    for i in range(n):
        ok = ok and _bufs(plans[i], 0, b, geo.t_in[i]) is not None and (i == 0 or _bufs(plans[i], 1, b, geo.t_in[i]) is not None)
    if ok:
        dummy = torch.empty(1, device="cuda")
        for i in range(1, n + 1):
            cons = plans[i] if i < n else None
            ok = ok and _link_ok(kind=0, b=b, prod=plans[i - 1], prod_t=geo.t_in[i - 1], cons=cons, cons_t=geo.t_in[i] if cons else 0, act=True,
                                 pool=pools[i - 1])
        for i in range(1, n):
            cons = plans[i - 1] if i - 1 >= 1 else None
            ok = ok and _link_ok(kind=2, b=b, prod=plans[i], prod_t=geo.t_in[i], cons=cons, cons_t=geo.t_in[i - 1] if cons else 0, act=True,
                                 pool=pools[i - 1], sact=dummy)
    spec = dict(plans=plans, geo=geo, pools=pools) if ok else None
    cache[key] = spec
    return spec


def encoder_forward(enc, x, needed=None):
    """All level outputs of Encoder.forward's conv stack ([pooled, activated] per level), or None (caller falls back).
    ``needed``: levels whose output the caller reads (None = all); under no_grad the others are not written (empty tensors)."""
    if not _enabled or ops._conv_impl == ops.IMPL_SIMT or x.requires_grad:
        return None
    spec = encoder_spec(enc, x.shape[0], x.shape[2])
    if spec is None:
        return None
    spec = dict(spec, needed=set(range(len(enc.convs))) if needed is None else set(needed), hook=wgrad_issued_hooks.get(enc),
                grad=torch.is_grad_enabled())
    ws = [c.weight for c in enc.convs]
    bs = [c.bias for c in enc.convs]
    return _EncoderStackFn.apply(spec, x, *ws, *bs)
