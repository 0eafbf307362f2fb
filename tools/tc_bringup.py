"""Bring-up check for the tcgen05 conv: each geometry runs in its own process (a hang only kills that process)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = {  # name: (level, ci, co, K, stride, T_in, upsample, unpool_level, B)
    "tiny_l3_s1": (3, 8, 16, 3, 1, 8, False, None, 4),
    "enc0": (0, 6, 12, 15, 2, 64, False, None, 32),
    "enc1": (1, 12, 24, 15, 2, 32, False, None, 32),
    "enc2": (2, 24, 48, 15, 2, 16, False, None, 32),
    "enc3": (3, 48, 96, 15, 2, 8, False, None, 32),
    "dec0": (3, 96, 48, 15, 1, 8, True, 3, 32),
    "dec1": (2, 48, 24, 15, 1, 16, True, 2, 32),
    "dec2": (1, 24, 12, 15, 1, 32, True, 1, 32),
    "dec3": (0, 24, 6, 15, 1, 64, True, 0, 32),
    "len8_enc0": (0, 6, 12, 3, 1, 8, False, None, 8),
    "traj3": (3, 24, 48, 31, 1, 128, False, None, 8),
    "enc0_b5": (0, 6, 12, 15, 2, 64, False, None, 5),
}


def run_case(name):
    import numpy as np
    import torch

    import hm_vae_b200 as H
    from hm_vae_b200 import ops

    lvl, ci, co, k, s, t_in, up, unpool_lvl, b = CASES[name]
    topo = json.load(open(os.path.join(ROOT, "tests", "golden", "topology.json")))["levels"]
    nb = topo[lvl]["neighbours"]
    j = len(nb)
    torch.manual_seed(1)
    conv = H.SkeletonConv(nb, j * ci, j * co, k, j, stride=s, padding=(k - 1) // 2, bias=True, padding_mode="reflection").cuda()
    kw = dict(lrelu=True)
    if unpool_lvl is not None:
        pl = topo[unpool_lvl]["pooling_list"]
        un = H.SkeletonUnpool(pl, ci)
        x = torch.randn(b, len(pl) * ci, t_in // 2, device="cuda")
        kw.update(upsample=True, unpool_src=un.src, src_joints=len(pl))
    else:
        x = torch.randn(b, j * ci, t_in, device="cuda")
    res = {}
    outs = {}
    for impl in (ops.IMPL_SIMT, ops.IMPL_TC):
        ops.set_conv_impl(impl)
        xi = x.clone().requires_grad_(True)
        conv.zero_grad()
        y = conv.fused_forward(xi, **kw)
        gy = torch.ones_like(y) if impl == ops.IMPL_SIMT and False else None
        torch.manual_seed(7)
        gy = torch.randn_like(y)
        y.backward(gy)
        torch.cuda.synchronize()
        outs[impl] = (y.detach(), xi.grad.detach())
    rel = lambda a, r: float((a - r).norm() / r.norm())
    res["y"] = rel(outs[ops.IMPL_TC][0], outs[ops.IMPL_SIMT][0])
    res["dx"] = rel(outs[ops.IMPL_TC][1], outs[ops.IMPL_SIMT][1])
    # same comparison without the LeakyReLU epilogue: isolates dgrad from sign flips of y ~ 0 between the two forwards
    kw2 = dict(kw, lrelu=False)
    o2 = {}
    for impl in (ops.IMPL_SIMT, ops.IMPL_TC):
        ops.set_conv_impl(impl)
        xi = x.clone().requires_grad_(True)
        conv.zero_grad()
        y = conv.fused_forward(xi, **kw2)
        torch.manual_seed(7)
        y.backward(torch.randn_like(y))
        torch.cuda.synchronize()
        o2[impl] = (xi.grad.detach(), conv.weight.grad.detach().clone(), conv.bias.grad.detach().clone())
    res["dx_nolrelu"] = rel(o2[ops.IMPL_TC][0], o2[ops.IMPL_SIMT][0])
    res["dw_nolrelu"] = rel(o2[ops.IMPL_TC][1], o2[ops.IMPL_SIMT][1])
    res["db_nolrelu"] = rel(o2[ops.IMPL_TC][2], o2[ops.IMPL_SIMT][2])
    res["dw_masked_zero"] = float((o2[ops.IMPL_TC][1] * (1 - conv.mask)).abs().sum()) == 0.0
    # timing
    ops.set_conv_impl(ops.IMPL_TC)
    for which in ("tc", "simt"):
        ops.set_conv_impl(ops.IMPL_TC if which == "tc" else ops.IMPL_SIMT)
        with torch.no_grad():
            for _ in range(3):
                conv.fused_forward(x, **kw)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                conv.fused_forward(x, **kw)
            e1.record()
            torch.cuda.synchronize()
            res["fwd_us_" + which] = e0.elapsed_time(e1) / 20 * 1e3
    nnz = sum(len(n) for n in nb)
    t_out = (t_in + 2 * ((k - 1) // 2) - k) // s + 1
    res["gflop"] = 2.0 * b * t_out * k * co * ci * nnz / 1e9
    res["tc_tflops"] = res["gflop"] / res["fwd_us_tc"] * 1e3 / 1e3
    print(name, json.dumps({k2: (round(v, 6) if isinstance(v, float) else v) for k2, v in res.items()}), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--one":
        run_case(sys.argv[2])
    else:
        names = sys.argv[1:] or list(CASES)
        for n in names:
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), "--one", n], capture_output=True, text=True, timeout=120)
                out = (r.stdout.strip().splitlines() or ["(no output)"])[-1]
                if r.returncode != 0:
                    out += " | rc=%d %s" % (r.returncode, r.stderr.strip().splitlines()[-1] if r.stderr.strip() else "")
                print(out, flush=True)
            except subprocess.TimeoutExpired:
                print(n, "TIMEOUT (hang)", flush=True)
                break
