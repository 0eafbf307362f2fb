"""Kernel timeline of ONE training step replayed from its CUDA graph (torch.profiler / CUPTI): start offset, duration, stream and
name of every kernel, in time order, plus per-stream busy time and the critical-path gaps on the main stream.
    python tools/timeline.py > gpurun_out/timeline.txt        (HMVAE_STACK=0: per-layer path; BATCH, CONFIG env)"""
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import yaml  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from hm_vae_b200 import ops, stack  # noqa: E402
from hm_vae_b200.trainer_motion_vae import Trainer  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
stack.set_enabled(os.environ.get("HMVAE_STACK", "1") != "0")
hp = yaml.safe_load(open(os.path.join(ROOT, "configs", os.environ.get("CONFIG", "len64_no_aug_hm_vae.yaml"))))
bs, T = int(os.environ.get("BATCH", "32")), hp["train_seq_len"]
torch.manual_seed(0)
tr = Trainer(dict(hp), device=dev, sync_losses=False).to(dev)
g = torch.Generator().manual_seed(1234)
rot = ops.rot6d_to_rotmat(torch.randn(bs, T, 24, 6, generator=g).to(dev))
data = (torch.stack((rot[..., 0], rot[..., 1]), dim=-2).reshape(bs, T, -1).contiguous(), rot.reshape(bs, T, -1).contiguous())
tr.enable_cuda_graph(data, hp, 0, warmup=3)
for _ in range(20):
    tr.gen_update(data, hp, 0)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        tr.gen_update(data, hp, 0)
    torch.cuda.synchronize()
path = os.path.join(tempfile.mkdtemp(), "trace.json")
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "dur" in e]
ev.sort(key=lambda e: e["ts"])
if not ev:
    print("no kernel events captured (CUPTI unavailable?)")
    sys.exit(1)
# split into steps at the optimiser clock tick (first kernel of a step's graph)
starts = [i for i, e in enumerate(ev) if "opt_clock_tick" in e["name"] or "clock_tick" in e["name"]]
if len(starts) >= 3:
    a, b = starts[1], starts[2]
else:
    a, b = 0, len(ev)
step = ev[a:b]
t0 = step[0]["ts"]
main = max(set(e["args"].get("stream", 0) for e in step), key=lambda s: sum(1 for e in step if e["args"].get("stream", 0) == s))
print("step: %d kernels, span %.1f us (next step starts at %.1f us), main stream %s" % (
    len(step), max(e["ts"] + e["dur"] for e in step) - t0, (ev[b]["ts"] - t0) if b < len(ev) else -1, main))
busy = {}
for e in step:
    s = e["args"].get("stream", 0)
    busy[s] = busy.get(s, 0.0) + e["dur"]
print("busy us per stream:", {k: round(v, 1) for k, v in busy.items()})
prev_end = None
for e in step:
    s = e["args"].get("stream", 0)
    name = e["name"].replace("hmvae::", "").replace("void ", "")
    name = name.split("(")[0][:60]
    gap = ""
    if s == main:
        if prev_end is not None:
            gap = "gap %5.1f" % (e["ts"] - prev_end)
        prev_end = e["ts"] + e["dur"]
    print("%8.1f  %6.1f  s%-3s %-10s %s" % (e["ts"] - t0, e["dur"], s, gap, name))
