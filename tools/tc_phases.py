"""Where does a conv_tc CTA spend its life?  %globaltimer stamps per CTA (hmvae_conv_tc_debug) for every len64 layer at B=32,
fprop and dgrad, warm (second call) -- prints, per layer: grid, kernel span, and medians of the phase lengths in microseconds:
  launch skew (first CTA entry -> this CTA's entry), prologue (entry -> barriers/TMEM/table ready), first stage wait,
  MMA issue span (first stage landed -> last MMA issued), drain (last issue -> accumulators complete), dump store, exit."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from hm_vae_b200 import ops  # noqa: E402
from hm_vae_b200._lib import lib  # noqa: E402
from oracle import topology as topo  # noqa: E402

torch.cuda.init()
dev = "cuda"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
levels = topo.hierarchy()
# (name, level, ci, co, stride, T_in, upsample, unpool)
layers = [("enc0", 0, 6, 12, 2, 64), ("enc1", 1, 12, 24, 2, 32), ("enc2", 2, 24, 48, 2, 16), ("enc3", 3, 48, 96, 2, 8),
          ("dec0", 3, 96, 48, 1, 8), ("dec1", 2, 48, 24, 1, 16), ("dec2", 1, 24, 12, 1, 32), ("dec3", 0, 24, 6, 1, 64)]
buf = torch.zeros(8 * 4096, dtype=torch.int64, device=dev)
for name, lvl, ci, co, s, T in layers:
    nb = levels[lvl]["neighbours"]
    J = len(nb)
    plan = ops.ConvPlan(nb, ci, co, 15, s, 7, "reflect")
    w = torch.randn(J * co, J * ci, 15, device=dev)
    x = torch.randn(B, J * ci, T, device=dev, requires_grad=True)
    for mode in ("fprop", "dgrad"):
        for rep in range(3):
            buf.zero_()
            if rep == 2:
                torch.cuda.synchronize()
                lib.hmvae_conv_tc_debug(buf.data_ptr())
            y = ops.skeleton_conv(x, w, None, plan)
            if mode == "dgrad":
                lib.hmvae_conv_tc_debug(None)
                gy = torch.randn_like(y)
                if rep == 2:
                    torch.cuda.synchronize()
                    buf.zero_()
                    lib.hmvae_conv_tc_debug(buf.data_ptr())
                (gx,) = torch.autograd.grad(y, x, gy)
            torch.cuda.synchronize()
            lib.hmvae_conv_tc_debug(None)
        t = buf.cpu().numpy().reshape(-1, 8).astype(np.float64)
        t = t[t[:, 0] > 0]
        t0 = t[:, 0].min()
        span = (t[:, 6].max() - t0) / 1e3
        med = lambda v: float(np.median(v)) / 1e3
        print("%s %-5s ctas %3d span %6.1f us | skew med %5.1f max %5.1f | prologue %4.1f | first stage %4.1f | mma issue %5.1f | drain %4.1f | "
              "dump %4.1f | exit %4.1f | cta life med %5.1f" % (
                  name, mode, len(t), span, med(t[:, 0] - t0), (t[:, 0].max() - t0) / 1e3, med(t[:, 1] - t[:, 0]), med(t[:, 2] - t[:, 1]),
                  med(t[:, 3] - t[:, 2]), med(t[:, 4] - t[:, 3]), med(t[:, 5] - t[:, 4]), med(t[:, 6] - t[:, 5]), med(t[:, 6] - t[:, 0])),
              flush=True)
