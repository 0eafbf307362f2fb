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


def workload_config(bs, world):
    """`config` of the JSON line -- IDENTICAL in both arms (the driver compares them)."""
    return {"workload": "configs/len64_no_aug_hm_vae.yaml train step (fwd+bwd+allreduce+Adam), B=%d per GPU, T=64, 24-joint SMPL, "
                        "random-init weights" % bs,
            "global_batch": world * bs, "parallelism": "dp%d" % world,
            "l2": "no flush: per-step working set (params+grads+Adam state ~265 MB) exceeds the 126 MB L2"}


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
    # same --steps / --warmup as the product arm (>= 3 warm-up like there); one step is ~0.15 s on 16 host cores, so the run is
    # bounded by capping the timed steps at 600 (~90 s)
    steps, warmup = min(args.steps, 600), max(args.warmup, 3)
    sec, threads = cpu_reference_step_time(hp, args.batch, steps, warmup)
    value = args.batch / sec
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(args.batch, 1),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": "%d timed steps of B=%d after %d warm-up (oracle/hmvae_ref.py, torch CPU fp32, fwd+bwd+Adam)" % (steps, args.batch, warmup)},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ reference on the same GPU
def reference_cuda_leg(hp, batch, dev, steps=30, warmup=5):
    """SURVEY 8(d) "reference timing alongside (i)": the reference's own CUDA-PyTorch path on this B200 -- eager torch ops
    (F.conv1d on weight*mask through cuDNN, matmul pool / unpool, nn.Upsample, the 23-step FK chain, autograd backward,
    torch.optim.Adam), cudnn.benchmark=True and torch's default TF32 policy as in train_motion_vae.py:16-18.  /root/reference
    does not exist on the GPU box, so this is the oracle port (op-for-op restatement, pinned against the real reference by
    tests/test_oracle_golden.py) moved to the device: kind = "port".  It is a measured baseline, never the product path."""
    from oracle import hmvae_ref as O

    was = torch.backends.cudnn.benchmark
    torch.backends.cudnn.benchmark = True
    try:
        d = np.load(os.path.join(ROOT, "hm_vae_b200", "data", "smpl24.npz"))
        parents, off = d["parents"].tolist(), torch.from_numpy(d["offsets"])
        ora = O.HMVAEOracle(hp, parents, off).init(seed=0)
        ora.set_params({k: v.detach() for k, v in ora.params.items()}, device=dev)
        opt = torch.optim.Adam(list(ora.params.values()), lr=hp["lr"], weight_decay=hp["weight_decay"])
        data = O.synthetic_batch(batch, hp["train_seq_len"], parents, off, seed=1234, device=dev)
        eps = O.draw_eps(ora, batch, seed=4321, device=dev)

        def step():
            opt.zero_grad(set_to_none=True)
            out = ora.step(data["seq_rot_6d"], data["seq_rot_mat"], eps, iterations=0)
            opt.step()
            return out

        for _ in range(warmup):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            out = step()
        e1.record()
        torch.cuda.synchronize()
        ms = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3) / steps
        return {"kind": "port", "what": "reference algorithm as eager CUDA-PyTorch ops on this GPU (oracle/hmvae_ref.py on cuda: cuDNN conv1d "
                                        "on weight*mask, autograd, torch.optim.Adam; cudnn.benchmark=True, default TF32 policy)",
                "ms_per_step": ms, "sequences_per_s": batch / (ms * 1e-3), "steps": steps, "warmup": warmup, "batch": batch,
                "loss": float(out["total"])}
    finally:
        torch.backends.cudnn.benchmark = was


def measure_tf32_peak(dev, n=8192, iters=10):
    """cuBLAS TF32 GEMM n^3 (allow_tf32), best of `iters` -- the conv roofline denominator ("to be measured")."""
    was = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        a = torch.randn(n, n, device=dev)
        b = torch.randn(n, n, device=dev)
        for _ in range(3):
            torch.matmul(a, b)
        best = None
        for _ in range(iters):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.matmul(a, b)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            best = ms if best is None else min(best, ms)
        del a, b
        return 2.0 * n ** 3 / (best * 1e-3) / 1e12
    finally:
        torch.backends.cuda.matmul.allow_tf32 = was


def other_configs_leg(dev):
    """BASELINE config 2 (len8, B=8: launch-overhead bound) and the config-5 trajectory model (T=128, K=31, B=8): training step
    through Trainer.gen_update as a CUDA graph (parity for both: tests/test_gpu_parity.py)."""
    from hm_vae_b200 import ops
    from hm_vae_b200.trainer_motion_vae import Trainer

    def timed(fn, steps):
        for _ in range(5):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    def hmvae_data(hp, bs):
        return synthetic_device_batch(bs, hp["train_seq_len"], dev, 1234)

    def traj_data(hp, bs):
        T = hp["train_seq_len"]
        g = torch.Generator().manual_seed(1234)
        d6, dm = hmvae_data(hp, bs)
        return (d6, dm, None, torch.randn(bs, T, 72, generator=g).to(dev), None, None, torch.randn(bs, T, 3, generator=g).to(dev))

    out = {}
    for tag, cfg_name, bs, make in (("len8", "len8_data_aug_hm_vae.yaml", 8, hmvae_data), ("trajectory", "trajectory_model.yaml", 8, traj_data)):
        try:
            hp = load_cfg(cfg_name)
            torch.manual_seed(0)
            tr = Trainer(dict(hp), device=dev, sync_losses=False).to(dev)
            data = make(hp, bs)
            tr.enable_cuda_graph(data, hp, 0, warmup=3)
            ms = timed(lambda: tr.gen_update(data, hp, 0), 100)
            res = tr.gen_update(data, hp, 0)
            out[tag] = {"config": "configs/" + cfg_name, "batch": bs, "ms_per_step": ms, "sequences_per_s": bs / (ms * 1e-3),
                        "launches_per_step": tr.launches_per_step, "loss": float(res[0]), "cuda_graph": True}
            tr.gen_opt.close()
            del tr
        except Exception as exc:        # reported, never fatal for the headline line
            out[tag] = {"error": "%s: %s" % (type(exc).__name__, exc)}
    # the latent-space optimisation loop of configs/len_64_test_interpolation.yaml (150 iterations: 51 on the
    # latents, 99 on the decoder copy; one window, bs = 1 as in the reference), two CUDA graphs, losses kept on the device
    try:
        from hm_vae_b200.seq_two_hier_sa_vae import TwoHierSAVAEModel

        hp = load_cfg("len_64_test_interpolation.yaml")
        torch.manual_seed(0)
        model = TwoHierSAVAEModel(dict(hp), device=dev).to(dev)
        d6, dm = synthetic_device_batch(1, hp["train_seq_len"], dev, 1234)
        T = hp["train_seq_len"]
        mask = torch.zeros(1, T, 24, device=dev)
        mask[:, ::hp["interpolation_window"]] = 1
        mask[:, -1] = 1
        model.optimize_latent(d6.view(1, T, 24, 6), dm.view(1, T, 24, 3, 3), mask, hp)          # warm-up (plans, allocator)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        res = model.optimize_latent(d6.view(1, T, 24, 6), dm.view(1, T, 24, 3, 3), mask, hp)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        hist = res["losses"].cpu()
        out["latent_optimisation"] = {"config": "configs/len_64_test_interpolation.yaml", "batch": 1, "iterations": int(hp["opt_it"]),
                                      "ms_total": ms, "ms_per_iteration": ms / hp["opt_it"], "loss_first": float(hist[0, 5]),
                                      "loss_last": float(hist[-1, 5]), "cuda_graph": True,
                                      "note": "includes graph capture of both phases (2 eager + 2 capture iterations)"}
        del model
    except Exception as exc:
        out["latent_optimisation"] = {"error": "%s: %s" % (type(exc).__name__, exc)}
    return out


# ------------------------------------------------------------------------------------------------ kernel instrumentation
class KernelTimer:
    """Wraps the C-ABI entry points with CUDA events (on the launching stream) to attribute step time to kernels."""

    CONV = {"fprop": ("hmvae_conv_fprop", "hmvae_conv_fprop_tc"), "dgrad": ("hmvae_conv_dgrad", "hmvae_conv_dgrad_tc"),
            "wgrad": ("hmvae_conv_wgrad", "hmvae_conv_wgrad_tc")}

    # the linked stack path (hm_vae_b200/stack.py) issues the phases of a conv separately: classified by their mode / kind argument
    STACK = ("hmvae_conv_tc_stage", "hmvae_conv_tc_run", "hmvae_conv_link")

    def __init__(self, lib, repeat=1):
        """repeat > 1: every conv compute entry point is issued `repeat` times back to back between its two events (the
        calls are idempotent), so that the launch queue is full and the event pair measures device time, not host gaps."""
        self.lib, self.records, self.orig, self.repeat = lib, [], {}, repeat
        self.conv_names = {n for v in self.CONV.values() for n in v} | set(self.STACK)

    @classmethod
    def klass(cls, name, a):
        """conv class of one recorded call: 'fprop' / 'dgrad' / 'wgrad' or None."""
        for c, names in cls.CONV.items():
            if name in names:
                return c
        if name in ("hmvae_conv_tc_stage", "hmvae_conv_tc_run"):
            return "fprop" if int(a[1]) == 0 else "dgrad"
        if name == "hmvae_conv_link":
            return "fprop" if int(a[0]._obj.kind) == 0 else "dgrad"
        return None

    def class_ms(self):
        """device ms per conv class over all recorded calls (stage + tensor-core kernel + finish / link)."""
        torch.cuda.synchronize()
        out = {c: 0.0 for c in self.CONV}
        for name, a, e0, e1, rep in self.records:
            c = self.klass(name, a)
            if c is not None:
                out[c] += e0.elapsed_time(e1) / rep
        return out

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
        rep = self.repeat if name in self.conv_names else 1

        def call(*a):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(rep):
                rc = fn(*a)
            e1.record()
            self.records.append((name, a, e0, e1, rep))
            return rc
        return call

    def __exit__(self, *exc):
        for name, fn in self.orig.items():
            setattr(self.lib, name, fn)

    def summary(self):
        torch.cuda.synchronize()
        out = {}
        for name, a, e0, e1, rep in self.records:
            d = out.setdefault(name, dict(ms=0.0, calls=0))
            d["ms"] += e0.elapsed_time(e1) / rep
            d["calls"] += 1
        return out

    def per_layer(self, steps):
        """conv time per (entry point, layer); layers are labelled by the order their plans first appear in a step."""
        torch.cuda.synchronize()
        order, out = {}, {}
        for name, a, e0, e1, rep in self.records:
            if not name.startswith("hmvae_conv_") or name in ("hmvae_conv_tc_supported", "hmvae_conv_tc_workspace", "hmvae_conv_packed_size"):
                continue
            if name == "hmvae_conv_link":
                key = a[0]._obj.prod                    # attributed to the producer conv of the boundary
            else:
                key = a[0].value if hasattr(a[0], "value") else a[0]
            idx = order.setdefault(key, len(order))
            label = name.replace("hmvae_conv_", "")
            if name in self.STACK:
                label = "%s_%s" % (self.klass(name, a), label)
            k = "%s[L%d]" % (label, idx)
            out[k] = out.get(k, 0.0) + e0.elapsed_time(e1) / rep / steps
        return {k: round(v * 1e3, 1) for k, v in sorted(out.items())}


def conv_flops(model, batch, enc_passes=1, dec_passes=1):
    """Algorithmic FLOPs of the unmasked blocks only: 2*B*T_out*K*co*ci*nnz per conv, forward."""
    enc = dec = 0
    ts = model.enc.timestep_list
    for i, conv in enumerate(model.enc.convs):
        nnz = sum(len(nb) for nb in conv.neighbour_list)
        enc += 2 * batch * ts[i + 1] * conv.kernel_size * conv.out_channels_per_joint * conv.in_channels_per_joint * nnz
    dts = model.dec.timestep_list
    for i, conv in enumerate(model.dec.convs):
        t_out = dts[i] * (2 if model.dec.upsample[i] else 1)
        nnz = sum(len(nb) for nb in conv.neighbour_list)
        dec += 2 * batch * t_out * conv.kernel_size * conv.out_channels_per_joint * conv.in_channels_per_joint * nnz
    return enc_passes * enc + dec_passes * dec


def conv_flops_per_step(model, batch):
    return conv_flops(model, batch)


def inference_leg(model, hp, dev, pk, batch=512, iters=10):
    """BASELINE config 4: len64 `test()` path at B=512 (1 encoder + 2 decoder passes + 2 rot6d + 3 FK, no grad) -- the
    regime where the conv kernels have enough rows to be tensor-pipe bound.  Reports sequences/s and the conv TFLOP/s
    (algorithmic FLOPs of 1 enc + 2 dec passes / device time spent inside the conv entry points)."""
    from hm_vae_b200 import _lib

    T = hp["train_seq_len"]
    d6, dm = synthetic_device_batch(batch, T, dev, 99)
    g = torch.Generator().manual_seed(5)
    ks = [len(p) for p in model.enc.pooling_list]
    lat = [model.shallow_latent_d] + [model.latent_d] * (len(ks) - 1)
    zs = [torch.randn(batch, k, d, generator=g).to(dev) for k, d in zip(ks, lat)]
    for _ in range(3):
        model.test((d6, dm), hp, 0, sampled_z_list=zs)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        model.test((d6, dm), hp, 0, sampled_z_list=zs)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    with KernelTimer(_lib.lib, repeat=4) as kt:
        for _ in range(2):
            model.test((d6, dm), hp, 0, sampled_z_list=zs)
    summ = kt.summary()
    conv_ms = kt.class_ms()["fprop"] / 2
    fl = conv_flops(model, batch, enc_passes=1, dec_passes=2)
    tf = fl / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else None
    return {"workload": "configs/len_64_test_interpolation.yaml test() path, B=%d, no grad, eager" % batch, "ms_per_call": ms,
            "sequences_per_s": batch / (ms * 1e-3), "conv_fprop_ms": conv_ms, "conv_algorithmic_gflop": fl / 1e9,
            "conv_tflops": tf, "conv_frac_of_tf32_peak": (tf / pk["tf32"]) if tf else None, "tf32_peak_tflops": pk["tf32"],
            "conv_us_per_layer": kt.per_layer(2)}


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


def fk_sweep(dev, pk, sizes=(43, 683, 10923, 174763, 699051), iters=20):
    """BASELINE config 3: FK and rot6d->R, fwd and bwd kernels timed through the C ABI (no autograd glue), achieved
    algorithmic GB/s (FK 48 / 84 B per joint-frame, rot6d 60 / 84).  The largest size is the headline."""
    from hm_vae_b200 import _lib
    from hm_vae_b200._lib import check, int_array, lib, ptr, stream
    from hm_vae_b200.fk_layer import load_smpl24

    parents, offsets, _ = load_smpl24()
    par = int_array(parents)
    off = torch.as_tensor(offsets, dtype=torch.float32, device=dev).contiguous()
    g = torch.Generator().manual_seed(1)
    out = {"sizes": {}}

    def timed(fn):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters * 1e-3

    for frames in sizes:
        x6 = torch.randn(frames, 24, 6, generator=g).to(dev)
        rot = torch.empty(frames, 24, 3, 3, device=dev)
        check(lib.hmvae_rot6d_fwd(ptr(x6), ptr(rot), frames * 24, stream()), "rot6d_fwd")
        gp = torch.randn(frames, 24, 3, generator=g).to(dev)
        gr = torch.randn(frames, 24, 3, 3, generator=g).to(dev)
        pos, grot, gx6 = torch.empty(frames, 24, 3, device=dev), torch.empty_like(rot), torch.empty_like(x6)
        jf = frames * 24
        t = {
            "fk_fwd": (48.0, timed(lambda: check(lib.hmvae_fk_fwd(ptr(rot), 9, ptr(off), None, par, 24, frames, ptr(pos), None, stream())))),
            "fk_bwd": (84.0, timed(lambda: check(lib.hmvae_fk_bwd(ptr(rot), 9, ptr(off), None, par, 24, frames, ptr(gp), None, ptr(grot), stream())))),
            "rot6d_fwd": (60.0, timed(lambda: check(lib.hmvae_rot6d_fwd(ptr(x6), ptr(rot), jf, stream())))),
            "rot6d_bwd": (84.0, timed(lambda: check(lib.hmvae_rot6d_bwd(ptr(x6), ptr(gr), ptr(gx6), jf, stream())))),
        }
        out["sizes"][str(frames)] = {k: {"us": round(sec * 1e6, 2), "gbs": round(bpj * jf / sec / 1e9, 1),
                                         "frac": round(bpj * jf / sec / 1e9 / pk["hbm"], 4)} for k, (bpj, sec) in t.items()}
        del x6, rot, gp, gr, pos, grot, gx6
    big = out["sizes"][str(sizes[-1])]
    out.update(frames=sizes[-1], fwd_gbs=big["fk_fwd"]["gbs"], bwd_gbs=big["fk_bwd"]["gbs"], fwd_frac=big["fk_fwd"]["frac"],
               bwd_frac=big["fk_bwd"]["frac"], peak_gbs=pk["hbm"],
               fwd_bwd_gbs=round((48.0 + 84.0) * sizes[-1] * 24 / ((big["fk_fwd"]["us"] + big["fk_bwd"]["us"]) * 1e-6) / 1e9, 1),
               note="kernel time through the C ABI, CUDA events, %d back-to-back launches; algorithmic bytes 48 B/jf fwd, 84 B/jf bwd; "
                    "the largest size moves 805 MB fwd / 1.4 GB bwd per launch (> 126 MB L2); the 43- and 683-frame sizes (1 K / 16 K "
                    "joint-frames) are launch-latency bound: read their `us`, not their GB/s" % iters)
    out["fwd_bwd_frac"] = round(out["fwd_bwd_gbs"] / pk["hbm"], 4)
    return out


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
    # TF32 dense peak: measured here (cuBLAS TF32 8192^3 burst) -- the conv roofline denominator
    pk["tf32"] = measure_tf32_peak(dev) if rank == 0 else pk["bf16"] / 2.0
    pk["tf32_src"] = "measured in this run: torch.matmul 8192^3 with allow_tf32 (cuBLAS TF32), best of 10, CUDA events"
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

    # ---- timed: inputs resident in HBM -- in the step's own input buffers when it is a CUDA graph (a device-side producer such as
    #      the batch assembly kernel writes there directly), so that no per-step copy sits in the timed region
    resident = trainer.static_inputs(data, hp, iters0) if args.graph else data
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        out = trainer.gen_update(resident, hp, iters0)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / args.steps
    loss_val = float(out[0])

    # ---- e2e: through the public call (Trainer.gen_update) with PINNED HOST inputs: every step copies its 2.95 MB of inputs
    #      host -> device inside the timed region and every step's five losses are read back to the host.  Like any input
    #      pipeline the loop is one step deep: the read-back of step i completes while step i+1 is already queued (two pinned
    #      result slots), which lets gen_update's copy stream overlap the next H2D with the running step.
    h6, hm = seq_rot_6d.cpu().pin_memory(), seq_rot_mat.cpu().pin_memory()
    dev_out = [torch.zeros(5, device=dev), torch.zeros(5, device=dev)]
    host_out = [torch.zeros(5, dtype=torch.float32).pin_memory() for _ in range(2)]
    out_ev = [None, None]
    for _ in range(3):
        trainer.gen_update((h6, hm), hp, iters0)
    barrier()
    t0 = time.perf_counter()
    e0.record()
    last_loss = None
    for i in range(args.steps):
        k = i & 1
        if out_ev[k] is not None:
            out_ev[k].synchronize()                      # the result of step i-2 has landed: consume it
            last_loss = float(host_out[k][0])
        out = trainer.gen_update((h6, hm), hp, iters0)
        torch.stack([o.reshape(()) for o in out[:5]], out=dev_out[k])
        host_out[k].copy_(dev_out[k], non_blocking=True)
        out_ev[k] = torch.cuda.Event()
        out_ev[k].record()
    for k in range(2):
        if out_ev[k] is not None:
            out_ev[k].synchronize()
            last_loss = float(host_out[k][0])
    e1.record()
    barrier()
    ms_e2e = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3) / args.steps
    t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])
    # The clock sampler (100 ms period) needs >= ~1 s under load: keep stepping (untimed) if the two timed regions were shorter.
    # The NUMBER of extra steps is derived from the all-reduced times, so every rank runs the same count (each step contains the
    # cross-rank optimiser kernel: ranks must make identical call sequences).
    n_extra = int(max(0.0, 1.2 - (ms + ms_e2e) * args.steps * 1e-3) / (ms * 1e-3)) + 1
    for _ in range(n_extra):
        trainer.gen_update(data, hp, iters0)
    barrier()
    clocks = sampler.stop()

    # ---- per-kernel attribution (eager, CUDA events around each C-ABI call, after the timed region).  The steps contain the
    #      gradient all-reduce, so EVERY rank runs them; only rank 0 reports.  Pass 1: one event pair per call (includes host
    #      launch gaps) for the step breakdown.  Pass 2: weight-gradient overlap off and every conv entry point issued 8x back
    #      to back inside its event pair -> device time per conv call, the roofline numerator's denominator.
    prof_steps = 5

    def eager_steps(kt_repeat):
        with KernelTimer(_lib.lib, repeat=kt_repeat) as kt:
            saved, trainer._graphs = trainer._graphs, {}
            sync_was = trainer._sync.enabled
            trainer._sync.enabled = True
            for _ in range(prof_steps):
                trainer.gen_update(data, hp, iters0)
            trainer._graphs = saved
            trainer._sync.enabled = sync_was
        barrier()
        return kt

    kt = eager_steps(1)
    ops._overlap["allowed"] = False
    kt8 = eager_steps(8)
    ops._overlap["allowed"] = True
    line = None
    if rank == 0:
        summ, summ8 = kt.summary(), kt8.summary()
        cls_ms = {c: v / prof_steps for c, v in kt8.class_ms().items()}
        fl_fwd = conv_flops_per_step(model, bs)
        top = max(cls_ms, key=lambda k: cls_ms[k])
        top_ms = cls_ms[top]
        achieved = fl_fwd / (top_ms * 1e-3) / 1e12 if top_ms > 0 else 0.0
        tf32_peak = pk["tf32"]
        total_kernel_ms = sum(v["ms"] for v in summ.values()) / prof_steps
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get("conv_" + top, {}).get("dram_bytes_per_launch_set")
        n_layers = len(model.enc.convs) + len(model.dec.convs)
        roofline = {"kernel": "hmvae_conv_%s%s (all %d layers of one step = one launch set)" % (top, "" if args.conv_impl == "simt" else "_tc", n_layers),
                    "bound": "tensor", "achieved": achieved, "peak": tf32_peak, "unit": "TFLOP/s",
                    "frac": achieved / tf32_peak, "traffic": traffic,
                    "peak_note": "TF32 dense peak %s (for scale: %s bf16 burst %.1f TF/s); conv math is %s.  At B=32 the layers "
                                 "are latency / weight-bandwidth bound (7.2 GFLOP and 41 MB of weights per pass): see conv_large_batch for "
                                 "the tensor-bound regime" % (
                        pk["tf32_src"], pk["src"], pk["bf16"],
                        "fp32 CUDA-core" if args.conv_impl == "simt" else "auto (tcgen05 TF32 where supported, else fp32 CUDA-core)"),
                    "algorithmic_flops_per_launch_set": fl_fwd,
                    "ms_per_launch_set": top_ms,
                    "conv_class_ms_per_step": {k: round(v, 4) for k, v in cls_ms.items()},
                    "conv_class_tflops": {k: round(fl_fwd / (v * 1e-3) / 1e12, 2) if v > 0 else None for k, v in cls_ms.items()},
                    "kernel_ms_per_step": {k: round(v["ms"] / prof_steps, 4) for k, v in sorted(summ.items(), key=lambda kv: -kv[1]["ms"])},
                    "kernel_ms_total_per_step_eager": total_kernel_ms,
                    "conv_us_per_layer": kt8.per_layer(prof_steps)}
        # the single-GPU legs (FK sweep, B=512 inference, CPU baseline) are reported at N = 1 only: at N > 1 the other ranks
        # would just wait for rank 0
        fk = fk_sweep(dev, pk) if (args.fk_sweep and world == 1) else None
        big = None
        if args.large_batch and world == 1:
            try:
                big = inference_leg(model, hp, dev, pk)
            except Exception as exc:        # reported, never fatal for the headline line
                big = {"error": "%s: %s" % (type(exc).__name__, exc)}
        cpu = None
        if args.cpu_baseline and world == 1:
            sec, threads = cpu_reference_step_time(hp, bs, 60, 3)
            cpu = {"value": bs / sec, "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": "60 timed steps of B=%d after 3 warm-up (oracle/hmvae_ref.py, torch CPU fp32, fwd+bwd+Adam)" % bs}
        ref_cuda = others = None
        if args.reference_cuda and world == 1:
            try:
                ref_cuda = reference_cuda_leg(hp, bs, dev)
                ref_cuda["speedup_value"] = (bs / (ms * 1e-3)) / ref_cuda["sequences_per_s"]
                ref_cuda["speedup_e2e"] = (bs / (ms_e2e * 1e-3)) / ref_cuda["sequences_per_s"]
            except Exception as exc:
                ref_cuda = {"error": "%s: %s" % (type(exc).__name__, exc)}
        if args.other_configs and world == 1:
            others = other_configs_leg(dev)
        cfg = workload_config(bs, world)
        line = {"metric": METRIC, "value": world * bs / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "tf32" if args.conv_impl != "simt" else "f32", "data": "synthetic",
                "config": cfg,
                "run": {"cuda_graph": bool(args.graph), "conv_impl": args.conv_impl, "data_parallel": trainer.dp_mode,
                        "dp_barrier_timed_out": bool(getattr(trainer.gen_opt, "timed_out", lambda: False)()),
                        "optimiser_arena_elems": int(getattr(trainer.gen_opt, "numel", 0)),
                        "optimiser_masked_elems_skipped": int(getattr(trainer.gen_opt, "masked_elems", 0))},
                "e2e": {"value": world * bs / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e,
                        "h2d_bytes_per_step": int(h6.numel() * 4 + hm.numel() * 4 + 8), "d2h_bytes_per_step": 20},
                "gpu_launches": int(launches_per_step * args.steps), "launches_per_step": int(launches_per_step),
                "clocks": clocks, "roofline": roofline, "fk": fk, "conv_large_batch": big, "cpu_baseline": cpu,
                "reference_cuda": ref_cuda, "other_configs": others, "loss": loss_val}
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
    ap.add_argument("--no-large-batch", dest="large_batch", action="store_false")
    ap.add_argument("--no-reference-cuda", dest="reference_cuda", action="store_false")
    ap.add_argument("--no-other-configs", dest="other_configs", action="store_false")
    args = ap.parse_args()
    hp = load_cfg("len64_no_aug_hm_vae.yaml")
    if args.impl == "reference":
        run_reference(args, hp)
    else:
        run_b200(args, hp)


if __name__ == "__main__":
    main()
