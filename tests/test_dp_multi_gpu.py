"""Multi-GPU parity of the fused data-parallel optimiser step (csrc/dp.cu), driver-visible: spawns W ranks with torchrun when
>= 2 GPUs are visible (skips otherwise) and runs
  * tools/dp_kernel_check.py -- kernel-level, tight bounds (reduced gradient 1e-6, moments 1e-5, parameters 5e-7, ranks
    bit-identical), on the unicast peer-load path AND on the NVSwitch multicast (multimem) path;
  * tools/dp_check.py        -- three Trainer steps of the len64 model, fused kernel vs NCCL all-reduce + multi-tensor Adam.
"""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _torchrun(script, world, env_extra, args=()):
    env = dict(os.environ, **env_extra)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tools", script)] + list(args)
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    return r.returncode, r.stdout + r.stderr


def _worlds():
    n = torch.cuda.device_count() if torch.cuda.is_available() else 0
    return [w for w in (2, 4, 8) if w <= n]


@pytest.mark.parametrize("multicast", ["0", "1"])
def test_dp_adam_kernel_multi_rank(multicast):
    worlds = _worlds()
    if not worlds:
        pytest.skip("needs >= 2 GPUs")
    for w in worlds:
        rc, out = _torchrun("dp_kernel_check.py", w, {"HMVAE_DP_MULTICAST": multicast})
        assert rc == 0, out[-4000:]
        assert out.count("PASS") == w, out[-4000:]
        if multicast == "1":
            assert "nvls_multicast" in out, "NVSwitch multicast mapping unavailable:\n" + out[-2000:]


@pytest.mark.parametrize("multicast", ["0", "1"])
def test_trainer_fused_dp_vs_nccl_multi_rank(multicast):
    worlds = _worlds()
    if not worlds:
        pytest.skip("needs >= 2 GPUs")
    w = worlds[-1] if multicast == "1" else worlds[0]
    rc, out = _torchrun("dp_check.py", w, {"HMVAE_DP_MULTICAST": multicast})
    assert rc == 0, out[-4000:]
