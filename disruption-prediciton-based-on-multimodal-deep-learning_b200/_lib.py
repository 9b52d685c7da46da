"""ctypes binding of libdp_b200.so (the C ABI declared in include/dp_b200.h).

The product path has no CPU or library fallback: if the shared library is missing or the
device is not sm_100, every entry point raises.  Only plumbing lives here (pointer
marshalling, error translation); all arithmetic is in csrc/*.cu.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdp_b200.so")
CSRC_DIR = os.path.join(_HERE, "csrc")

DP_OK = 0
DP_F32, DP_BF16 = 0, 1
IMPL_AUTO, IMPL_SIMT, IMPL_TC = 0, 1, 2
LOSS_CE, LOSS_FOCAL, LOSS_LDAM = 0, 1, 2
DP_MAX_PARTS = 592

_ERR_NAMES = {-1: "DP_ERR_SHAPE", -2: "DP_ERR_ALIGN", -3: "DP_ERR_ARCH", -4: "DP_ERR_CUDA", -5: "DP_ERR_UNSUPPORTED"}


class ConvDesc(C.Structure):
    """Mirror of `dp_conv_desc` (include/dp_b200.h)."""

    _fields_ = [(n, C.c_int32) for n in (
        "B", "Ti", "Hi", "Wi", "C", "Cp", "To", "Ho", "Wo", "K", "Kp",
        "kt", "kh", "kw", "st", "sh", "sw", "pt", "ph", "pw", "dtype")]


class BnFin(C.Structure):
    """Mirror of `dp_bn_fin` (include/dp_b200.h): BatchNorm finalisation run by the last CTA of the kernel that produces
    the partial sums.  kind 1 = forward statistics, kind 2 = backward sums."""

    _fields_ = [("kind", C.c_int32), ("C", C.c_int32), ("Cp", C.c_int32), ("coef_zero", C.c_int32), ("count", C.c_double),
                ("gamma", C.c_void_p), ("beta", C.c_void_p), ("eps", C.c_float), ("momentum", C.c_float),
                ("running_mean", C.c_void_p), ("running_var", C.c_void_p),
                ("mean", C.c_void_p), ("rstd", C.c_void_p), ("scale", C.c_void_p), ("shift", C.c_void_p),
                ("dgamma", C.c_void_p), ("dbeta", C.c_void_p), ("coef", C.c_void_p), ("ticket", C.c_void_p)]


_vp, _i, _i64, _f, _d, _sz = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_double, C.c_size_t
_pdesc = C.POINTER(ConvDesc)
_pfin = C.POINTER(BnFin)
_pint = C.POINTER(C.c_int)

# name -> (restype, argtypes); exactly the declarations of include/dp_b200.h
SIGNATURES = {
    "dp_version": (_i, []),
    "dp_last_error": (C.c_char_p, []),
    "dp_device_check": (_i, []),
    "dp_num_sms": (_i, []),
    "dp_launch_count": (C.c_ulonglong, []),
    "dp_simt_launch_count": (C.c_ulonglong, []),
    "dp_simt_fallback_count": (C.c_ulonglong, []),
    "dp_set_option": (_i, [C.c_char_p, _i]),
    "dp_get_option": (_i, [C.c_char_p]),
    "dp_set_debug_buffer": (_i, [_vp, _sz]),
    "dp_conv_describe_plan": (_i, [_pdesc, _i, _i, C.c_char_p, _sz]),
    "dp_ncdhw_f32_to_ndhwc": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "dp_ndhwc_to_ncdhw_f32": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "dp_u8_frames_to_ndhwc": (_i, [_vp, _vp, C.POINTER(C.c_float), _i, _i, _i, _i, _i, _i, _vp]),
    "dp_pack_weights": (_i, [_pdesc, _vp, _vp, _vp, _vp]),
    "dp_conv_supported": (_i, [_pdesc, _i, _i]),
    "dp_conv_fwd": (_i, [_pdesc, _vp, _vp, _vp, _vp, _pint, _i, _vp]),
    "dp_conv_fwd_bnact": (_i, [_pdesc, C.POINTER(C.c_longlong), _vp, _vp, _vp, _f, _vp, _f, _vp, _i, _vp]),
    "dp_conv_dgrad": (_i, [_pdesc, _vp, _vp, _vp, _vp, _i, _vp]),
    "dp_dgrad_classes_weight_elems": (_sz, [_pdesc, _i]),
    "dp_pack_weights_dgrad_classes": (_i, [_pdesc, _vp, _vp, _vp]),
    "dp_conv_dgrad_classes": (_i, [_pdesc, _vp, _vp, _vp, _vp, _vp]),
    "dp_conv_dgrad_bnstats": (_i, [_pdesc, _vp, _vp, _vp, _vp, _vp, _vp, _f, _vp, _pint, _i, _vp]),
    "dp_conv_fwd_fin": (_i, [_pdesc, _vp, _vp, _vp, _vp, _pfin, _i, _vp]),
    "dp_conv_dgrad_bnstats_fin": (_i, [_pdesc, _vp, _vp, _vp, _vp, _vp, _vp, _f, _vp, _pfin, _i, _vp]),
    "dp_conv_wgrad_workspace": (_sz, [_pdesc, _i]),
    "dp_conv_wgrad": (_i, [_pdesc, _vp, _vp, _vp, _vp, _sz, _i, _vp]),
    "dp_stem_supported": (_i, [_pdesc]),
    "dp_stem_input_elems": (_sz, [_pdesc]),
    "dp_stem_pack_input_f32": (_i, [_pdesc, _vp, _vp, _vp]),
    "dp_stem_pack_input_u8": (_i, [_pdesc, _vp, C.POINTER(C.c_float), _vp, _vp]),
    "dp_stem_weight_elems": (_sz, [_pdesc]),
    "dp_stem_pack_weights": (_i, [_pdesc, _vp, _vp, _vp]),
    "dp_stem_conv_fwd": (_i, [_pdesc, _vp, _vp, _vp, _vp, _pint, _vp]),
    "dp_stem_conv_fwd_fin": (_i, [_pdesc, _vp, _vp, _vp, _vp, _pfin, _vp]),
    "dp_stem_conv_fwd_bnact": (_i, [_pdesc, _vp, _vp, _vp, _f, _vp, _vp]),
    "dp_stem_wgrad_workspace": (_sz, [_pdesc]),
    "dp_stem_conv_wgrad": (_i, [_pdesc, _vp, _vp, _vp, _vp, _sz, _vp]),
    "dp_bn_stats": (_i, [_vp, _i64, _i, _i, _vp, _pint, _vp]),
    "dp_bn_finalize": (_i, [_vp, _i, _i, _i, _d, _vp, _vp, _f, _f, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "dp_bn_eval_coeffs": (_i, [_vp, _vp, _vp, _vp, _f, _i, _i, _vp, _vp, _vp]),
    "dp_bn_act_apply": (_i, [_vp, _vp, _vp, _f, _vp, _f, _vp, _i64, _i, _i, _vp]),
    "dp_bn_act_bwd_reduce": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _f, _f, _vp, _pint, _i64, _i, _i, _vp]),
    "dp_bn_act_bwd_reduce_fin": (_i, [_vp, _vp, _vp, _vp, _vp, _f, _f, _vp, _i64, _i, _i, _pfin, _vp]),
    "dp_bn_bwd_finalize": (_i, [_vp, _i, _i, _i, _d, _vp, _vp, _vp, _vp, _vp, _vp]),
    "dp_bn_act_bwd_apply": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _f, _f, _vp, _vp, _i64, _i, _i, _vp]),
    "dp_add": (_i, [_vp, _vp, _vp, _i64, _i, _vp]),
    "dp_avgpool_fwd": (_i, [_vp, _vp, _i, _i64, _i, _i, _i, _vp]),
    "dp_avgpool_bwd": (_i, [_vp, _vp, _i, _i64, _i, _i, _i, _vp]),
    "dp_se_swish_fwd": (_i, [_vp, _vp, _vp, _i, _i64, _i, _i, _i, _vp]),
    "dp_se_swish_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i64, _i, _i, _i, _vp]),
    "dp_maxpool_hw_fwd": (_i, [_vp, _vp, _vp, _i64, _i, _i, _i, _i, _vp]),
    "dp_maxpool_hw_bwd": (_i, [_vp, _vp, _vp, _i64, _i, _i, _i, _i, _vp]),
    "dp_concat_channels": (_i, [_vp, _vp, _vp, _i64, _i, _i, _i, _i, _i, _i, _vp]),
    "dp_split_channels": (_i, [_vp, _vp, _vp, _i64, _i, _i, _i, _i, _i, _i, _vp]),
    "dp_loss_workspace": (_sz, [_i64]),
    "dp_loss_fwd_bwd": (_i, [_i, _vp, _vp, _vp, _vp, _f, _f, _i64, _i, _vp, _vp, _vp, _vp]),
    "dp_loss_bwd_scale": (_i, [_vp, _vp, _vp, _vp, _i64, _vp]),
    "dp_optim_workspace": (_sz, [_i64]),
    "dp_grad_sqnorm": (_i, [_vp, _i64, _f, _i, _vp, _vp, _vp]),
    "dp_adamw_apply": (_i, [_vp, _vp, _vp, _vp, _i64, _f, _f, _f, _f, _f, _i, _f, _f, _vp, _vp, _vp]),
    "dp_clip_adamw_step": (_i, [_vp, _vp, _vp, _vp, _i64, _f, _f, _f, _f, _f, _i, _f, _f, _vp, _vp, _vp]),
}

_lib = None
_lock = threading.Lock()
_device_ok = False


class DpError(RuntimeError):
    pass


def build(force: bool = False) -> str:
    """Compile csrc/*.cu for sm_100a into libdp_b200.so (in-tree)."""
    if force:
        subprocess.run(["make", "-C", CSRC_DIR, "clean"], check=True, stdout=subprocess.DEVNULL)
    subprocess.run(["make", "-C", CSRC_DIR, "-j", str(min(16, os.cpu_count() or 4))], check=True,
                   stdout=subprocess.DEVNULL)
    return LIB_PATH


def load() -> C.CDLL:
    """Load the shared library and attach the prototypes.  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise DpError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                f"(or `make -C {CSRC_DIR}`). There is no CPU or library fallback on this path.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError here == header/library mismatch
            fn.restype = res
            fn.argtypes = args
        # DP_OPTIONS="pdl=0,tc_mt=1": planner / launch options (dp_set_option) for A/B measurements
        for item in filter(None, os.environ.get("DP_OPTIONS", "").split(",")):
            name, _, value = item.partition("=")
            if lib.dp_set_option(name.strip().encode(), int(value)) != DP_OK:
                raise DpError(f"DP_OPTIONS: unknown option '{name}'")
        _lib = lib
    return _lib


def last_error() -> str:
    msg = load().dp_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(rc: int, what: str = "") -> None:
    if rc != DP_OK:
        raise DpError(f"{what or 'dp_b200'} failed with {_ERR_NAMES.get(rc, rc)}: {last_error()}")


def require_device() -> None:
    """Fail loudly unless the current CUDA device can run the sm_100a kernels."""
    global _device_ok
    if _device_ok:
        return
    import torch

    if not torch.cuda.is_available():
        raise DpError("dp_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback on this path")
    check(load().dp_device_check(), "dp_device_check")
    _device_ok = True


def stream_ptr() -> int:
    import torch

    return torch.cuda.current_stream().cuda_stream


def set_option(name: str, value: int) -> None:
    check(load().dp_set_option(name.encode(), int(value)), f"dp_set_option({name})")


def get_option(name: str) -> int:
    return int(load().dp_get_option(name.encode()))
