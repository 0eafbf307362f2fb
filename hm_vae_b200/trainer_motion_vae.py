"""Host mirror of the reference Trainer (trainer_motion_vae.py:15-135, 251-283) around the B200 step.

Same surface: ``Trainer(cfg)``, ``gen_update(data, hp, iterations, multigpus, validation_flag)``, ``save`` /
``resume`` / ``load_ckpt`` with the reference's file names and dict layout (``gen_%08d.pt`` = {'state_dict': ...},
``optimizer.pt`` = {'gen': ...}), Adam(lr, weight_decay) + StepLR.  Differences:
  * the optimiser is the fused multi-tensor kernel (``ops.FusedAdam``), LR schedule evaluated on the host;
  * the 9 ``.item()`` host syncs per step of the reference (:95-98) are opt-in (``sync_losses``);
  * ``enable_cuda_graph`` captures forward+backward+all-reduce+Adam once and replays it each step;
  * multi-GPU is one process per GPU with bucketed NCCL all-reduce (``ddp.BucketedAllReduce``), not DataParallel.
"""
import math
import os
import sys

import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.init as init

from . import ops
from .ddp import BucketedAllReduce
from .seq_two_hier_sa_vae import TwoHierSAVAEModel
from .trajectory_pred_model import TrajectoryModel


def weights_init(init_type='gaussian'):
    """trainer_motion_vae.py:264-283 -- only modules whose class name STARTS with 'Conv' or 'Linear' are touched,
    i.e. nn.Linear but not SkeletonConv."""

    def init_fun(m):
        classname = m.__class__.__name__
        if (classname.find('Conv') == 0 or classname.find('Linear') == 0) and hasattr(m, 'weight'):
            if init_type == 'gaussian':
                init.normal_(m.weight.data, 0.0, 0.02)
            elif init_type == 'xavier':
                init.xavier_normal_(m.weight.data, gain=math.sqrt(2))
            elif init_type == 'kaiming':
                init.kaiming_normal_(m.weight.data, a=0, mode='fan_in')
            elif init_type == 'orthogonal':
                init.orthogonal_(m.weight.data, gain=math.sqrt(2))
            elif init_type == 'default':
                pass
            else:
                assert 0, "Unsupported initialization: {}".format(init_type)
            if hasattr(m, 'bias') and m.bias is not None:
                init.constant_(m.bias.data, 0.0)

    return init_fun


def get_model_list(dirname, key):
    if not os.path.exists(dirname):
        return None
    models = sorted(os.path.join(dirname, f) for f in os.listdir(dirname)
                    if os.path.isfile(os.path.join(dirname, f)) and key in f and ".pt" in f)
    return models[-1] if models else None


class _NoCollective:
    """Stand-in for BucketedAllReduce when the collective is inside the optimiser kernel (dp_fused)."""
    fused = True

    def __init__(self, world):
        self.world, self.enabled = world, True

    def begin(self):
        pass

    def finish(self):
        pass

    def allreduce_now(self):
        pass


class Trainer(nn.Module):
    def __init__(self, cfg, device=None, sync_losses=True, n_buckets=4, dp_fused=None):
        super(Trainer, self).__init__()
        if cfg['model_name'] == "TrajectoryModel":
            self.model = TrajectoryModel(cfg, device=device)
        elif cfg['model_name'] == "TwoHierSAVAEModel":
            self.model = TwoHierSAVAEModel(cfg, device=device)
        else:
            raise ValueError("unknown model_name %r" % cfg['model_name'])
        self.cfg = cfg
        self.sync_losses = sync_losses
        self.apply(weights_init(cfg['init']))
        self.base_lr = cfg['lr']
        self.gen_opt = None
        self._n_buckets = n_buckets
        # optimiser / data parallelism: None = flat-arena Adam whose kernel is also the collective (peer-memory reduce-scatter
        # + Adam + all-gather when world > 1), falling back to NCCL bucketed all-reduce + multi-tensor Adam if peer memory cannot
        # be mapped (or HMVAE_DP_FUSED=0); True / False force one or the other
        self._dp_fused = dp_fused
        self.dp_mode = None
        self._sync = None
        self._graphs = {}
        self._static = None

    # ------------------------------------------------------------------ optimiser / schedule
    def _ensure_opt(self):
        if self.gen_opt is None:
            params = [p for p in self.model.parameters() if p.requires_grad]
            world = dist.get_world_size() if dist.is_initialized() else 1
            fused = self._dp_fused
            if fused is None:
                fused = os.environ.get("HMVAE_DP_FUSED", "1") != "0"
            if fused:
                try:
                    from .dp_fused import FusedDataParallelAdam
                    from .skeleton import SkeletonConv
                    # structurally dead weight entries (skeleton.py:84-96: zero at init, multiplied by the mask in forward):
                    # the fused step leaves them out of its reduce-scatter / Adam / all-gather sweep
                    masks = {id(mod.weight): mod.mask for mod in self.model.modules() if isinstance(mod, SkeletonConv)}
                    self.gen_opt = FusedDataParallelAdam(params, lr=self.base_lr, weight_decay=self.cfg['weight_decay'],
                                                         prefer=os.environ.get("HMVAE_DP_PEER", "symm"), masks=masks)
                    self._configure_schedule()
                    self._sync = _NoCollective(world)
                    self.dp_mode = "fused_peer_memory(%s)" % self.gen_opt.arenas.backend
                    # Buckets in reverse-layer order: the decoder's share of the optimiser step (and of the collective) starts as
                    # soon as the decoder's gradients are final and runs under the encoder's backward.  Measured (ms per step,
                    # graph replay, A/B inside one box): 1 GPU 0.754 -> 0.742 with a second bucket for the deepest encoder level;
                    # 8 GPUs on the NVSwitch multicast path 0.861 -> 0.831 with the decoder bucket on 32 CTAs x 4 loads in flight
                    # (a second bucket there: 0.910 -- its flag barrier lands where the encoder backward has no slack);
                    # 2 GPUs (unicast peer loads, 4x the bytes per rank) 0.768 -> 0.805: off.  profiles/r02_dp_overlap_ab_*.log
                    dec = getattr(self.model, "dec", None)
                    split = os.environ.get("HMVAE_DP_SPLIT", "auto")
                    # across ranks only where a rank's share of the bucket is small enough to fit the window: 8 ranks on the
                    # multicast path (4 ranks, same settings: 0.844 -> 0.886 ms, off)
                    nvls = bool(self.gen_opt.arenas.mc_grad) and world >= int(os.environ.get("HMVAE_DP_SPLIT_MIN_WORLD", "8"))
                    if dec is not None and ((world == 1 or nvls) if split == "auto" else split != "0"):
                        enc_ids = {id(p) for p in self.model.enc.parameters()}
                        dec_ids = {id(p) for p in dec.parameters()} - enc_ids
                        self.model.mid_backward = lambda: self.gen_opt.step_partial(dec_ids, grad_scale=1.0 / world)
                        self.dp_mode += "+split"
                        # second bucket (reverse-layer order): the deepest encoder conv -- a quarter of the arena at len64 -- and
                        # the encoder's latent heads, as soon as that level's weight-gradient kernels have been issued; the
                        # step's tail then only covers the shallow encoder levels
                        enc_bucket = os.environ.get("HMVAE_DP_SPLIT_ENC", "auto")
                        if (world == 1 if enc_bucket == "auto" else enc_bucket != "0") and hasattr(self.model.enc, "convs"):
                            import weakref
                            from . import stack
                            convs = list(self.model.enc.convs)
                            conv_ids = {id(p) for c in convs for p in c.parameters()}
                            top_ids = {id(p) for p in convs[-1].parameters() if p.requires_grad}
                            bucket_ids = top_ids | (enc_ids - conv_ids)
                            me = weakref.ref(self)

                            def _bucket(layer, n_layers):
                                tr = me()
                                if tr is not None and tr.gen_opt is not None and layer == n_layers - 1:
                                    # the conv's .grad is assigned only when the stack's backward node returns: vouch for it
                                    tr.gen_opt.step_partial(bucket_ids, grad_scale=1.0 / world, written=top_ids)
                            stack.wgrad_issued_hooks[self.model.enc] = _bucket
                            self.dp_mode += "2"
                    return
                except ops._lib.HmvaeError as exc:
                    if self._dp_fused:
                        raise
                    print("hm_vae_b200: %s -- falling back to the NCCL all-reduce path" % exc, file=sys.stderr)
            self.gen_opt = ops.FusedAdam(params, lr=self.base_lr, weight_decay=self.cfg['weight_decay'])
            self._configure_schedule()
            self._sync = BucketedAllReduce(self.model, n_buckets=self._n_buckets)
            self.dp_mode = "nccl_bucketed_allreduce" if world > 1 else "single"

    def _configure_schedule(self):
        """StepLR (trainer_motion_vae.py:251-262) is evaluated on the device from the optimiser's own iteration counter: like
        the reference's scheduler.step(), it advances once per gen_update call."""
        if self.cfg.get('lr_policy', 'constant') == 'step':
            self.gen_opt.set_schedule(self.cfg['gamma'], self.cfg['step_size'])

    def lr_at(self, iterations):
        """StepLR(step_size, gamma) (trainer_motion_vae.py:251-262); 'constant' when no policy is configured."""
        if self.cfg.get('lr_policy', 'constant') == 'step':
            return self.base_lr * self.cfg['gamma'] ** (iterations // self.cfg['step_size'])
        return self.base_lr

    # ------------------------------------------------------------------ one step
    def _device_step(self, data, hp, iterations):
        self.gen_opt.zero_grad(set_to_none=True)
        if self._sync.fused:
            self.gen_opt.begin_step()
        self._sync.begin()
        out = self.model(data, hp, iterations)
        self._sync.finish()
        self.gen_opt.step_dyn(grad_scale=1.0 / self._sync.world)
        return out

    def _fwd_bwd(self, data, hp, iterations):
        """forward + backward only, no collective (the part that is captured as a CUDA graph when world > 1)."""
        self.gen_opt.zero_grad(set_to_none=True)
        return self.model(data, hp, iterations)

    def _record(self, out, validation):
        names = ["total", "kl", "rec_6d", "rec_rot", "rec_pose", "rec_joint_pos", "rec_root_v", "rec_linear_v", "rec_angular_v"]
        prefix = "loss_val_" if validation else "loss_"
        # every loss is already a scalar (the reference's torch.mean over DataParallel replicas is the identity here)
        sc = lambda v: v.reshape(()) if v.numel() == 1 else torch.mean(v)
        for n, v in zip(names, out[:9]):
            setattr(self, prefix + n, sc(v))
        if len(out) > 9:
            for i, v in enumerate(out[9]):
                setattr(self, "loss_hier_kl_%d" % (i + 1), sc(v))
        if not self.sync_losses:
            return tuple(sc(v) for v in out[:9])
        vals = torch.stack([sc(v) for v in out[:9]]).tolist()      # ONE device->host sync
        self._check_health()
        return tuple(vals)

    def _check_health(self):
        """Raises if the fused data-parallel kernel ever gave up waiting for a peer (needs a host sync: called where the host
        synchronises anyway, in save(), and every 500 steps otherwise)."""
        chk = getattr(self.gen_opt, "check_health", None)
        if chk is not None:
            chk()

    def gen_update(self, data, hp, iterations, multigpus=False, validation_flag=False):
        self._ensure_opt()
        if validation_flag:
            with torch.no_grad():
                out = self.model(data, hp, iterations, validation_flag=True)
            return self._record(out, True)
        key = self._graph_key(hp, iterations, data)
        self._steps = getattr(self, "_steps", 0) + 1
        if not self.sync_losses and self._steps % 500 == 0:
            self._check_health()
        self.gen_opt.advance()
        if key in self._graphs:
            graph, static_in, static_out = self._graphs[key]
            self._load_inputs(static_in, data)
            graph.replay()
            out = static_out
            if self._sync.world > 1 and not self._sync.fused:
                # NCCL path: the graph holds forward+backward; the NCCL all-reduce and the Adam kernel follow eagerly
                self._sync.allreduce_now()
                self.gen_opt.step_dyn(grad_scale=1.0 / self._sync.world)
        else:
            out = self._device_step(data, hp, iterations)
        return self._record(out, False)

    def _load_inputs(self, static_in, data):
        """Moves this step's inputs into the graph's static buffers.  Host (pinned) inputs go through a copy stream into a
        double-buffered staging area first, so the H2D transfer of step i+1 runs while the graph of step i is still executing
        (the static buffers themselves are read until the loss kernel at the end of the forward pass)."""
        host = [i for i, (dst, src) in enumerate(zip(static_in, data)) if dst is not None and torch.is_tensor(src) and not src.is_cuda]
        if not host:
            for dst, src in zip(static_in, data):
                if dst is not None and src.data_ptr() != dst.data_ptr():     # the caller may hand back the static buffers themselves
                    dst.copy_(src, non_blocking=True)
            return
        if getattr(self, "_stage", None) is None:
            self._copy_stream = torch.cuda.Stream()
            self._stage = [[torch.empty_like(t) if t is not None else None for t in static_in] for _ in range(2)]
            self._stage_ev = [None, None]
            self._stage_free = [None, None]
            self._stage_k = 0
        k = self._stage_k = self._stage_k ^ 1
        main = torch.cuda.current_stream()
        with torch.cuda.stream(self._copy_stream):
            if self._stage_free[k] is not None:
                self._copy_stream.wait_event(self._stage_free[k])        # the step that used this slot has consumed it
            for i in host:
                self._stage[k][i].copy_(data[i], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
        main.wait_event(ev)
        for i, (dst, src) in enumerate(zip(static_in, data)):
            if dst is None:
                continue
            dst.copy_(self._stage[k][i] if i in host else src, non_blocking=True)
        free = torch.cuda.Event()
        free.record()
        self._stage_free[k] = free

    # ------------------------------------------------------------------ CUDA graph of the whole step
    def _graph_key(self, hp, iterations, data):
        detach = iterations < hp.get('iteration_interval', 0)
        shapes = tuple(tuple(d.shape) if torch.is_tensor(d) else None for d in data)
        return (detach, shapes)

    def static_inputs(self, data, hp, iterations):
        """The captured step's own input buffers for this batch shape (after ``enable_cuda_graph``): a producer that writes the next
        batch straight into them (e.g. the on-device batch assembly) and passes them to ``gen_update`` avoids the per-step copy."""
        return tuple(self._graphs[self._graph_key(hp, iterations, data)][1])

    def enable_cuda_graph(self, data, hp, iterations, warmup=3):
        """Captures fwd + bwd (+ all-reduce) + Adam for this batch shape / detach phase.  Inputs are copied into static
        device buffers before each replay; the LR and Adam's bias corrections reach the graph through a pinned-host
        -> device copy node (``FusedAdam.step_dyn``).  The warm-up steps are real optimisation steps."""
        self._ensure_opt()
        dev = next(self.model.parameters()).device
        static_in = [d.to(device=dev, dtype=torch.float32).clone() if torch.is_tensor(d) else None for d in data]
        multi = self._sync.world > 1 and not self._sync.fused      # the fused peer-memory step has no NCCL call: one graph
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self.gen_opt.advance()
                self._device_step(static_in, hp, iterations)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.gen_opt.zero_grad(set_to_none=True)
        graph = torch.cuda.CUDAGraph()
        n0 = ops._lib.launch_count()
        ops._force_repack = True            # the packed tf32 weight copies must be refreshed inside every replay
        self._sync.enabled = not multi      # no NCCL inside the graph: with world > 1 only forward+backward is captured
        try:
            cap = torch.cuda.Stream(priority=ops.main_stream_priority())
            with torch.cuda.graph(graph, stream=cap):
                static_out = self._fwd_bwd(static_in, hp, iterations) if multi else self._device_step(static_in, hp, iterations)
        finally:
            ops._force_repack = False
        if multi:
            # the capture pass itself executed nothing: leave the hooks off for good (eager all-reduce after each replay)
            self._sync.enabled = False
        self.launches_per_step = ops._lib.launch_count() - n0
        self._graphs[self._graph_key(hp, iterations, data)] = (graph, static_in, static_out)
        return graph

    # ------------------------------------------------------------------ checkpoints (reference layout)
    def save(self, snapshot_dir, iterations, multigpus=False):
        """Reference file names and dict layouts (trainer_motion_vae.py:108-113): gen_%08d.pt = {'state_dict': model.state_dict()},
        optimizer.pt = {'gen': torch.optim.Adam-format state_dict}.  With world > 1 this is COLLECTIVE: every rank must call it
        (the Adam moments are sharded across ranks); only rank 0 writes, and all ranks leave together."""
        self._ensure_opt()
        self._check_health()
        world = dist.get_world_size() if dist.is_initialized() else 1
        rank = dist.get_rank() if dist.is_initialized() else 0
        opt_state = self.gen_opt.state_dict()             # gathers the sharded moments (collective)
        if rank == 0:
            gen_name = os.path.join(snapshot_dir, 'gen_%08d.pt' % (iterations + 1))
            opt_name = os.path.join(snapshot_dir, 'optimizer.pt')
            torch.save({'state_dict': self.model.state_dict()}, gen_name, _use_new_zipfile_serialization=False)
            torch.save({'gen': opt_state}, opt_name, _use_new_zipfile_serialization=False)
        if world > 1:
            dist.barrier()

    def resume(self, checkpoint_dir, hp, multigpus=False):
        self._ensure_opt()
        last_model_name = get_model_list(checkpoint_dir, "gen")
        self.load_ckpt(last_model_name)
        iterations = int(last_model_name[-11:-3])
        state = torch.load(os.path.join(checkpoint_dir, 'optimizer.pt'), weights_only=False)
        self.gen_opt.load_state_dict(state['gen'])
        # the reference re-creates StepLR with last_epoch = iterations (trainer_motion_vae.py:126): same schedule position here
        self.gen_opt.set_clock(self.gen_opt.step_count, iterations)
        print('Resume from iteration %d' % iterations)
        return iterations

    def load_ckpt(self, ckpt_name):
        state_dict = torch.load(ckpt_name, weights_only=False)
        model_dict = self.model.state_dict()
        model_dict.update(state_dict['state_dict'])
        self.model.load_state_dict(model_dict)

    def test(self, data, hp, iterations):
        return self.model.test(data, hp, iterations)
