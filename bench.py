#!/usr/bin/env python
"""hm-vae hot-path benchmark (driver contract: one JSON line on rank 0).

    python bench.py --gpus 1 --steps 50 --warmup 10
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...
    python bench.py --impl reference            # the reference algorithm (oracle port, torch CPU, all host threads)

Workload: configs/len64_no_aug_hm_vae.yaml training step (fwd + bwd + gradient all-reduce + Adam), B=32 sequences per
GPU (BASELINE config 1/5 shape), 24-joint SMPL, synthetic rotations, random-init weights.  metric = sequences/s.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import yaml  # noqa: E402

METRIC, UNIT = "train_sequences_per_sec", "sequences/s"


def load_cfg(name):
    return yaml.safe_load(open(os.path.join(ROOT, "configs", name)))


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], bf16=d["bf16_tflops"], bf16_sustained=d.get("bf16_tflops_sustained"), src="measured")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, src="fallback")


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
                for n, v in zip(names, f[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm))


# ------------------------------------------------------------------------------------------------ CPU baseline (oracle)
def cpu_reference_step_time(hp, batch, steps, warmup):
    """The reference algorithm on the host cores: oracle port (torch CPU fp32), fwd + bwd + torch.optim.Adam."""
    from oracle import hmvae_ref as O

    torch.set_num_threads(os.cpu_count() or 1)
    d = np.load(os.path.join(ROOT, "hm_vae_b200", "data", "smpl24.npz"))
    parents, off = d["parents"].tolist(), torch.from_numpy(d["offsets"])
    ora = O.HMVAEOracle(hp, parents, off).init(seed=0)
    opt = torch.optim.Adam(list(ora.params.values()), lr=hp["lr"], weight_decay=hp["weight_decay"])
    data = O.synthetic_batch(batch, hp["train_seq_len"], parents, off, seed=1234)
    eps = O.draw_eps(ora, batch, seed=4321)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        ora.step(data["seq_rot_6d"], data["seq_rot_mat"], eps, iterations=0)
        opt.step()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    return float(np.mean(times)), torch.get_num_threads()


def run_reference(args, hp):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = min(args.steps, 20), min(max(args.warmup, 1), 3)
    sec, threads = cpu_reference_step_time(hp, args.batch, steps, warmup)
    value = args.batch / sec
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "len64_no_aug_hm_vae train step (fwd+bwd+Adam), B=%d, T=64, 24-joint SMPL" % args.batch},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": "%d timed steps of B=%d after %d warm-up (oracle/hmvae_ref.py, torch CPU fp32)" % (steps, args.batch, warmup)},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ kernel instrumentation
class KernelTimer:
    """Wraps the C-ABI entry points with CUDA events (on the launching stream) to attribute step time to kernels."""

    CONV = ("hmvae_conv_fprop", "hmvae_conv_dgrad", "hmvae_conv_wgrad")

    def __init__(self, lib):
        self.lib, self.records, self.orig = lib, [], {}

    def __enter__(self):
        from hm_vae_b200 import _lib

        for name in _lib.EXPORTS:
            if name in ("hmvae_last_error", "hmvae_version", "hmvae_launch_count", "hmvae_conv_plan_create", "hmvae_conv_plan_destroy"):
                continue
            fn = getattr(self.lib, name)
            self.orig[name] = fn
            setattr(self.lib, name, self._wrap(name, fn))
        return self

    def _wrap(self, name, fn):
        def call(*a):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = fn(*a)
            e1.record()
            self.records.append((name, a, e0, e1))
            return rc
        return call

    def __exit__(self, *exc):
        for name, fn in self.orig.items():
            setattr(self.lib, name, fn)

    def summary(self):
        torch.cuda.synchronize()
        out = {}
        for name, a, e0, e1 in self.records:
            d = out.setdefault(name, dict(ms=0.0, calls=0))
            d["ms"] += e0.elapsed_time(e1)
            d["calls"] += 1
        return out

    def per_layer(self, steps):
        """conv time per (entry point, layer); layers are labelled by the order their plans first appear in a step."""
        torch.cuda.synchronize()
        order, out = {}, {}
        for name, a, e0, e1 in self.records:
            if not name.startswith("hmvae_conv_") or name in ("hmvae_conv_tc_supported", "hmvae_conv_tc_workspace", "hmvae_conv_packed_size"):
                continue
            key = a[0].value if hasattr(a[0], "value") else a[0]
            idx = order.setdefault(key, len(order))
            k = "%s[L%d]" % (name.replace("hmvae_conv_", ""), idx)
            out[k] = out.get(k, 0.0) + e0.elapsed_time(e1) / steps
        return {k: round(v * 1e3, 1) for k, v in sorted(out.items())}


def conv_flops_per_step(model, batch):
    """Algorithmic FLOPs of the unmasked blocks only: 2*B*T_out*K*co*ci*nnz per conv, fwd; x3 for fwd+bwd."""
    total = 0
    t = model.max_timesteps
    ts = model.enc.timestep_list
    for i, conv in enumerate(model.enc.convs):
        nnz = sum(len(nb) for nb in conv.neighbour_list)
        total += 2 * batch * ts[i + 1] * conv.kernel_size * conv.out_channels_per_joint * conv.in_channels_per_joint * nnz
    dts = model.dec.timestep_list
    for i, conv in enumerate(model.dec.convs):
        t_out = dts[i] * (2 if model.dec.upsample[i] else 1)
        nnz = sum(len(nb) for nb in conv.neighbour_list)
        total += 2 * batch * t_out * conv.kernel_size * conv.out_channels_per_joint * conv.in_channels_per_joint * nnz
    return total


# ------------------------------------------------------------------------------------------------ main arm
def synthetic_device_batch(bs, t, dev, seed):
    """SURVEY 8d inputs, built with the product's own kernels: x6 ~ N(0,1) -> R -> (6D, rotmat)."""
    from hm_vae_b200 import ops

    g = torch.Generator(device="cpu").manual_seed(seed)
    x6 = torch.randn(bs, t, 24, 6, generator=g).to(dev)
    rot = ops.rot6d_to_rotmat(x6)
    seq_rot_mat = rot.reshape(bs, t, -1).contiguous()
    seq_rot_6d = torch.stack((rot[..., 0], rot[..., 1]), dim=-2).reshape(bs, t, -1).contiguous()
    return seq_rot_6d, seq_rot_mat


def fk_sweep(dev, pk, frames=699051, iters=10):
    """BASELINE config 3 at its largest size: FK fwd and bwd, achieved algorithmic GB/s."""
    import hm_vae_b200 as H

    fk = H.ForwardKinematicsLayer(device=dev)
    g = torch.Generator().manual_seed(1)
    rot = H.rotation_matrix_from_ortho6d(torch.randn(frames, 24, 6, generator=g).to(dev)).requires_grad_(True)
    gp = torch.randn(frames, 24, 3, generator=g).to(dev)
    res = {}
    for _ in range(3):
        pos = fk(rot)
        pos.backward(gp)
        rot.grad = None
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    tf = tb = 0.0
    for _ in range(iters):
        e[0].record()
        pos = fk(rot)
        e[1].record()
        pos.backward(gp)
        e[2].record()
        torch.cuda.synchronize()
        tf += e[0].elapsed_time(e[1])
        tb += e[1].elapsed_time(e[2])
        rot.grad = None
    jf = frames * 24
    res["frames"] = frames
    res["fwd_gbs"] = 48.0 * jf / (tf / iters * 1e-3) / 1e9
    res["bwd_gbs"] = 84.0 * jf / (tb / iters * 1e-3) / 1e9
    res["fwd_frac"] = res["fwd_gbs"] / pk["hbm"]
    res["bwd_frac"] = res["bwd_gbs"] / pk["hbm"]
    res["note"] = "algorithmic bytes 48 B/jf fwd, 84 B/jf bwd (bwd time includes autograd glue); inputs 604 MB > L2"
    return res


def run_b200(args, hp):
    import torch.distributed as dist

    from hm_vae_b200 import _lib, ddp, ops
    from hm_vae_b200.trainer_motion_vae import Trainer

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback (use --impl reference)")
    rank, world, local = ddp.init_from_env("nccl")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    pk = peaks()
    ops.set_conv_impl({"auto": 0, "simt": 1, "tc": 2}[args.conv_impl])

    torch.manual_seed(0)
    trainer = Trainer(dict(hp), device=dev, sync_losses=False).to(dev)
    model = trainer.model
    ddp.broadcast_parameters(model)
    bs, T = args.batch, hp["train_seq_len"]
    seq_rot_6d, seq_rot_mat = synthetic_device_batch(bs, T, dev, 1234 + rank)
    data = (seq_rot_6d, seq_rot_mat)
    iters0 = 0

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (real optimisation steps), optional whole-step CUDA graph
    if args.graph:
        trainer.enable_cuda_graph(data, hp, iters0, warmup=max(args.warmup, 3))
        launches_per_step = trainer.launches_per_step
        trainer.gen_update(data, hp, iters0)
    else:
        n0 = _lib.launch_count()
        for _ in range(max(args.warmup, 3)):
            trainer.gen_update(data, hp, iters0)
        torch.cuda.synchronize()
        launches_per_step = (_lib.launch_count() - n0) // max(args.warmup, 3)

    # ---- timed: inputs resident in HBM
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        out = trainer.gen_update(data, hp, iters0)
    e1.record()
    barrier()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1) / args.steps
    loss_val = float(out[0])

    # ---- e2e: pinned host inputs, H2D inside the timed region, D2H of the step's losses every step
    h6, hm = seq_rot_6d.cpu().pin_memory(), seq_rot_mat.cpu().pin_memory()
    host_out = torch.zeros(5, dtype=torch.float32).pin_memory()
    for _ in range(3):
        trainer.gen_update((h6, hm), hp, iters0)
    barrier()
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        out = trainer.gen_update((h6, hm), hp, iters0)
        host_out.copy_(torch.stack([o.reshape(()) for o in out[:5]]), non_blocking=False)
    e1.record()
    barrier()
    ms_e2e = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3) / args.steps

    t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])

    # ---- per-kernel attribution (eager, CUDA events around each C-ABI call, after the timed region).  The steps contain the
    #      gradient all-reduce, so EVERY rank runs them; only rank 0 reports.
    prof_steps = 5
    with KernelTimer(_lib.lib) as kt:
        saved, trainer._graphs = trainer._graphs, {}
        sync_was = trainer._sync.enabled
        trainer._sync.enabled = True
        for _ in range(prof_steps):
            trainer.gen_update(data, hp, iters0)
        trainer._graphs = saved
        trainer._sync.enabled = sync_was
    barrier()
    line = None
    if rank == 0:
        summ = kt.summary()
        conv = {k: summ.get(k, dict(ms=0.0, calls=0)) for k in KernelTimer.CONV}
        fl_fwd = conv_flops_per_step(model, bs)
        top = max(conv, key=lambda k: conv[k]["ms"])
        top_ms = conv[top]["ms"] / prof_steps
        achieved = fl_fwd / (top_ms * 1e-3) / 1e12 if top_ms > 0 else 0.0
        tf32_peak = pk["bf16"] / 2.0
        total_kernel_ms = sum(v["ms"] for v in summ.values()) / prof_steps
        roofline = {"kernel": top, "bound": "tensor", "achieved": achieved, "peak": tf32_peak, "unit": "TFLOP/s",
                    "frac": achieved / tf32_peak, "traffic": None,
                    "peak_note": "TF32 dense taken as 1/2 of the %s bf16 figure (%.1f TF/s); conv math is %s" % (
                        pk["src"], pk["bf16"], "fp32 CUDA-core" if args.conv_impl == "simt" else "auto (tcgen05 TF32 where supported, else fp32 CUDA-core)"),
                    "algorithmic_flops_per_launch_set": fl_fwd,
                    "ms_per_step_in_kernel": top_ms,
                    "kernel_ms_per_step": {k: round(v["ms"] / prof_steps, 4) for k, v in sorted(summ.items(), key=lambda kv: -kv[1]["ms"])},
                    "kernel_ms_total_per_step_eager": total_kernel_ms,
                    "conv_us_per_layer": kt.per_layer(prof_steps)}
        fk = fk_sweep(dev, pk) if args.fk_sweep else None
        cpu = None
        if args.cpu_baseline:
            sec, threads = cpu_reference_step_time(hp, bs, 10, 2)
            cpu = {"value": bs / sec, "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": "10 timed steps of B=%d after 2 warm-up (oracle/hmvae_ref.py, torch CPU fp32, fwd+bwd+Adam)" % bs}
        line = {"metric": METRIC, "value": world * bs / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "tf32" if args.conv_impl != "simt" else "f32", "data": "synthetic",
                "config": {"workload": "configs/len64_no_aug_hm_vae.yaml train step (fwd+bwd+allreduce+Adam), B=%d per GPU, "
                                       "T=64, 24-joint SMPL, random-init weights" % bs,
                           "global_batch": world * bs, "parallelism": "dp%d" % world, "cuda_graph": bool(args.graph),
                           "conv_impl": args.conv_impl,
                           "l2": "no flush: per-step working set (params+grads+Adam state ~265 MB) exceeds the 126 MB L2"},
                "e2e": {"value": world * bs / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e,
                        "h2d_bytes_per_step": int(h6.numel() * 4 + hm.numel() * 4 + 8), "d2h_bytes_per_step": 20},
                "gpu_launches": int(launches_per_step * args.steps), "launches_per_step": int(launches_per_step),
                "clocks": clocks, "roofline": roofline, "fk": fk, "cpu_baseline": cpu, "loss": loss_val}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(line), flush=True)


def main():
    import faulthandler
    faulthandler.dump_traceback_later(600, exit=True)      # a hung collective prints where every thread is and exits
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="sequences per GPU")
    ap.add_argument("--conv-impl", default="auto", choices=["auto", "simt", "tc"])
    ap.add_argument("--no-graph", dest="graph", action="store_false")
    ap.add_argument("--no-cpu-baseline", dest="cpu_baseline", action="store_false")
    ap.add_argument("--no-fk-sweep", dest="fk_sweep", action="store_false")
    args = ap.parse_args()
    hp = load_cfg("len64_no_aug_hm_vae.yaml")
    if args.impl == "reference":
        run_reference(args, hp)
    else:
        run_b200(args, hp)


if __name__ == "__main__":
    main()
