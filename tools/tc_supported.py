"""Prints which (layer geometry, batch, T) the tcgen05 conv kernels accept: fprop / dgrad / wgrad.  Geometries: all conv layers of
the trajectory model (K=31, T=128) and of the len64 / len8 HM-VAE."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from hm_vae_b200 import ops  # noqa: E402
from hm_vae_b200._lib import lib  # noqa: E402
from oracle import topology as topo  # noqa: E402

torch.cuda.init()
levels = topo.hierarchy()
cases = [("traj L%d" % i, i, 3 * 2 ** i, 6 * 2 ** i, 31, 1, 8, 128) for i in range(4)]
cases += [("traj L%d B=32" % i, i, 3 * 2 ** i, 6 * 2 ** i, 31, 1, 32, 128) for i in range(4)]
cases += [("T=256 L0", 0, 3, 6, 31, 1, 4, 256), ("T=200 L1", 1, 6, 12, 15, 1, 4, 200), ("T=300 s2 L0", 0, 6, 12, 15, 2, 4, 300)]
for name, lvl, ci, co, k, s, b, t in cases:
    plan = ops.ConvPlan(levels[lvl]["neighbours"], ci, co, k, s, (k - 1) // 2, "reflect")
    print("%-16s ci=%d co=%d K=%d s=%d B=%d T=%d: fprop %d dgrad %d wgrad %d" % (
        name, ci, co, k, s, b, t, lib.hmvae_conv_tc_supported(plan.handle, b, t, 0), lib.hmvae_conv_tc_supported(plan.handle, b, t, 1),
        lib.hmvae_conv_wgrad_tc_supported(plan.handle, b, t)))
