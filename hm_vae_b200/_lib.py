"""ctypes binding of libhmvae_b200.so (the C ABI in include/hmvae_b200.h).

The library is the product: if it is missing or cannot be loaded, importing this module raises -- there is no
PyTorch/CPU fallback anywhere in the package.
"""
import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_double, c_float, c_int, c_long, c_longlong, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libhmvae_b200.so")


class HmvaeError(RuntimeError):
    pass


def _build_library(force=False):
    if os.environ.get("HMVAE_AUTOBUILD", "1") != "1":
        raise HmvaeError("libhmvae_b200.so is missing or stale: run `python -m hm_vae_b200.build` (nvcc, sm_100a)")
    import importlib.util

    import fcntl

    spec = importlib.util.spec_from_file_location("_hmvae_build", os.path.join(_HERE, "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    # one builder at a time: the ranks of a torchrun job import this module concurrently
    with open(os.path.join(_HERE, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if force or _stale():
                mod.build(force=force)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _stale():
    """Library older than a kernel source / header (dlopen caches by path, so this is decided BEFORE loading)."""
    if not os.path.exists(LIB_PATH):
        return True
    import shutil

    if shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"):
        return False
    t = os.path.getmtime(LIB_PATH)
    csrc = os.path.join(_HERE, "csrc")
    deps = [os.path.join(csrc, f) for f in os.listdir(csrc)] + [os.path.join(os.path.dirname(_HERE), "include", "hmvae_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


if _stale():
    _build_library()


class ConvDesc(Structure):
    _fields_ = [(n, c_int) for n in ("joints", "ci", "co", "ksize", "stride", "pad", "pad_mode", "upsample", "src_joints",
                                     "lrelu", "out_joint_stride", "out_chan_offset", "out_channels_last")]


class AdamTensor(Structure):
    _fields_ = [("p", c_void_p), ("g", c_void_p), ("m", c_void_p), ("v", c_void_p), ("numel", c_long)]


class ConvLinkDesc(Structure):
    _fields_ = [("kind", c_int), ("batch", c_int), ("prod", c_void_p), ("prod_t", c_int), ("cons", c_void_p), ("cons_t", c_int),
                ("act", c_int), ("pool_joints", c_int), ("pool_off", POINTER(c_int)), ("pool_idx", POINTER(c_int)),
                ("dump", c_void_p), ("bias", c_void_p), ("aux", c_void_p), ("add", c_void_p), ("sact", c_void_p),
                ("yact_c", c_void_p), ("s_out", c_void_p), ("stage_ws", c_void_p)]


class HeadLevel(Structure):
    _fields_ = [("rows", c_int), ("features", c_int), ("d", c_int), ("kl_scale", c_float), ("x", c_void_p), ("enc_w", c_void_p),
                ("enc_b", c_void_p), ("eps", c_void_p), ("dec_w", c_void_p), ("dec_b", c_void_p), ("dist", c_void_p), ("z", c_void_p),
                ("feat", c_void_p), ("kl_acc", c_void_p), ("gfeat", c_void_p), ("gdist", c_void_p), ("gx", c_void_p), ("gfeat_stride", c_long)]


class RegTensor(Structure):
    _fields_ = [("p", c_void_p), ("p0", c_void_p), ("g", c_void_p), ("numel", c_long), ("accumulate", c_int)]


DP_MAX_WORLD, DP_MAX_RANGES = 8, 64


class DpPeers(Structure):
    _fields_ = [("world", c_int), ("rank", c_int), ("grad", c_void_p * DP_MAX_WORLD), ("param", c_void_p * DP_MAX_WORLD),
                ("flags", c_void_p * DP_MAX_WORLD), ("mc_grad", c_void_p), ("mc_param", c_void_p)]


P = c_void_p
IP = POINTER(c_int)
_SIGS = {
    "hmvae_last_error": (c_char_p, []),
    "hmvae_version": (c_int, []),
    "hmvae_launch_count": (c_longlong, []),
    "hmvae_conv_plan_create": (c_int, [POINTER(ConvDesc), IP, IP, IP, POINTER(c_void_p)]),
    "hmvae_conv_plan_destroy": (None, [c_void_p]),
    "hmvae_conv_tc_debug": (c_int, [P]),
    "hmvae_conv_tc_sizes": (c_int, [c_void_p, c_int, c_int, c_int, POINTER(c_long), POINTER(c_long)]),
    "hmvae_conv_tc_stage": (c_int, [c_void_p, c_int, P, P, c_int, c_int, P, P]),
    "hmvae_conv_tc_run": (c_int, [c_void_p, c_int, P, c_int, c_int, P, P, P]),
    "hmvae_conv_tc_finish": (c_int, [c_void_p, c_int, P, P, P, c_int, c_int, P]),
    "hmvae_conv_link_supported": (c_int, [POINTER(ConvLinkDesc)]),
    "hmvae_conv_link": (c_int, [POINTER(ConvLinkDesc), P]),
    "hmvae_conv_fprop": (c_int, [P, P, P, P, P, c_int, c_int, c_int, P]),
    "hmvae_conv_dgrad": (c_int, [P, P, P, P, P, c_int, c_int, c_int, P]),
    "hmvae_conv_wgrad": (c_int, [P, P, P, P, P, P, c_int, c_int, c_int, c_int, P]),
    "hmvae_conv_prologue_bwd": (c_int, [P, P, P, P, c_int, c_int, P]),
    "hmvae_conv_tc_supported": (c_int, [P, c_int, c_int, c_int]),
    "hmvae_conv_packed_size": (c_int, [P, POINTER(c_long), POINTER(c_long)]),
    "hmvae_conv_pack_weights": (c_int, [P, P, P, P, P]),
    "hmvae_conv_tc_workspace": (c_long, [P, c_int, c_int, c_int]),
    "hmvae_conv_fprop_tc": (c_int, [P, P, P, P, P, c_int, c_int, P, c_long, P]),
    "hmvae_conv_dgrad_tc": (c_int, [P, P, P, P, P, c_int, c_int, P, c_long, P]),
    "hmvae_conv_wgrad_tc_supported": (c_int, [P, c_int, c_int]),
    "hmvae_conv_wgrad_tc_workspace": (c_long, [P, c_int, c_int]),
    "hmvae_conv_wgrad_tc": (c_int, [P, P, P, P, P, P, c_int, c_int, c_int, P, c_long, P]),
    "hmvae_conv_wgrad_tc_stage_x": (c_int, [P, P, c_int, c_int, P, c_long, P]),
    "hmvae_pool_fwd": (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, IP, IP, c_int, P]),
    "hmvae_pool_bwd": (c_int, [P, P, P, c_int, c_int, c_int, c_int, c_int, IP, IP, c_int, P]),
    "hmvae_unpool_fwd": (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, IP, P]),
    "hmvae_unpool_bwd": (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, IP, P]),
    "hmvae_upsample2_fwd": (c_int, [P, P, c_long, c_int, P]),
    "hmvae_upsample2_bwd": (c_int, [P, P, c_long, c_int, P]),
    "hmvae_lrelu_fwd": (c_int, [P, P, c_long, c_float, P]),
    "hmvae_lrelu_bwd": (c_int, [P, P, P, c_long, c_float, P]),
    "hmvae_transpose_ct": (c_int, [P, P, c_int, c_int, c_int, P]),
    "hmvae_fk_fwd": (c_int, [P, c_int, P, P, IP, c_int, c_long, P, P, P]),
    "hmvae_fk_bwd": (c_int, [P, c_int, P, P, IP, c_int, c_long, P, P, P, P]),
    "hmvae_rot6d_fwd": (c_int, [P, P, c_long, P]),
    "hmvae_rot6d_bwd": (c_int, [P, P, P, c_long, P]),
    "hmvae_aa2rot_fwd": (c_int, [P, P, c_long, P]),
    "hmvae_latent_fwd": (c_int, [P, P, P, P, c_long, c_int, P]),
    "hmvae_latent_bwd": (c_int, [P, P, P, P, P, c_long, c_int, c_float, P]),
    "hmvae_recon_fwdbwd": (c_int, [P, c_int, P, P, P, IP, c_int, c_int, c_int, c_float, c_float, c_float, P, P, P, P, P]),
    "hmvae_recon_masked_fwdbwd": (c_int, [P, c_int, P, P, P, P, IP, c_int, c_int, c_int, c_float, c_float, c_float, P, P, P, P, P, P]),
    "hmvae_l2_reg_fwdbwd": (c_int, [POINTER(RegTensor), c_int, c_float, P, P]),
    "hmvae_latent_heads_fwd": (c_int, [POINTER(HeadLevel), c_int, P]),
    "hmvae_latent_heads_bwd": (c_int, [POINTER(HeadLevel), c_int, P]),
    "hmvae_linear_fwd": (c_int, [P, P, P, P, c_int, c_int, c_int, P]),
    "hmvae_linear_bwd": (c_int, [P, P, P, P, P, P, c_int, c_int, c_int, P]),
    "hmvae_loss_finalize": (c_int, [P, P, POINTER(c_float), POINTER(c_float), POINTER(c_float), c_int, P]),
    "hmvae_mse_fwd": (c_int, [P, P, P, c_long, P]),
    "hmvae_mse_bwd": (c_int, [P, P, P, c_long, c_float, P]),
    "hmvae_traj_fwdbwd": (c_int, [P, P, POINTER(c_float), POINTER(c_float), c_int, c_int, c_int, c_float, c_float, P, P, P]),
    "hmvae_adam_step": (c_int, [POINTER(AdamTensor), c_int, c_float, c_double, c_double, c_float, c_float, c_int, c_float, P]),
    "hmvae_adam_step_dyn": (c_int, [POINTER(AdamTensor), c_int, P, c_double, c_double, c_float, c_float, c_float, P]),
    "hmvae_opt_clock_tick": (c_int, [P, c_float, c_float, c_int, c_double, c_double, P, P]),
    "hmvae_rand_rotation": (c_int, [P, ctypes.c_double, P, c_long, P]),
    "hmvae_batch_assemble": (c_int, [P, P, P, P, c_int, c_int, P, P, P, P, P, P, P, P]),
    "hmvae_dp_adam_step": (c_int, [POINTER(DpPeers), P, P, POINTER(c_long), c_int, P, c_double, c_double, c_float, c_float, c_float, P, c_int, P]),
    "hmvae_dp_adam_step_units": (c_int, [POINTER(DpPeers), P, P, P, c_long, P, c_double, c_double, c_float, c_float, c_float, P, c_int, c_int, P]),
    "hmvae_ipc_alloc": (c_int, [c_long, POINTER(c_void_p)]),
    "hmvae_ipc_free": (c_int, [c_void_p]),
    "hmvae_ipc_get_handle": (c_int, [c_void_p, P]),
    "hmvae_ipc_open_handle": (c_int, [P, POINTER(c_void_p)]),
    "hmvae_ipc_close_handle": (c_int, [c_void_p]),
}
EXPORTS = sorted(_SIGS)


def _bind():
    handle = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in _SIGS.items():
        fn = getattr(handle, name)     # AttributeError here == the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    return handle


try:
    lib = _bind()
except AttributeError:                  # stale library from an older source tree: rebuild once, then fail loudly
    _build_library(force=True)
    lib = _bind()


def check(rc, what=""):
    if rc != 0:
        msg = lib.hmvae_last_error()
        raise HmvaeError("%s failed (code %d): %s" % (what or "hmvae call", rc, msg.decode() if msg else "?"))


def int_array(values):
    return (c_int * len(values))(*[int(v) for v in values])


def ptr(t):
    """Device pointer of a contiguous fp32 CUDA tensor (None -> NULL).  CPU tensors are an error: no fallback."""
    if t is None:
        return None
    if not t.is_cuda:
        raise HmvaeError("hm_vae_b200 ops run on CUDA tensors only (got a %s tensor); there is no CPU fallback" % t.device)
    if t.dtype != torch.float32:
        raise HmvaeError("hm_vae_b200 ops need float32 tensors (got %s)" % t.dtype)
    if not t.is_contiguous():
        raise HmvaeError("hm_vae_b200 ops need contiguous tensors")
    return t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def aligned(t):
    """Contiguous fp32 view with a 16-byte aligned base pointer (clones only when needed)."""
    t = t.contiguous()
    if t.data_ptr() % 16:
        t = t.clone()
    return t


def launch_count():
    return int(lib.hmvae_launch_count())
