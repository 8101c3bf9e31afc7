"""ctypes binding of libftb.so (the C ABI declared in include/ftb.h).

The library is built in-tree by ``__graft_entry__.build()``.  There is no fallback: if the
shared object is missing or a symbol is absent, importing the package fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FTB_LIB_PATH") or os.path.join(_HERE, "csrc", "libftb.so")   # override: A/B builds
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "ftb.h")

FTB_MAX_STAGES = 8


class FtbUnetCfg(C.Structure):
    _fields_ = [
        ("dim", C.c_int),
        ("n_stages", C.c_int),
        ("dim_mults", C.c_int * FTB_MAX_STAGES),
        ("data_channels", C.c_int),
        ("time_resolution", C.c_int),
        ("attn_heads", C.c_int),
        ("attn_dim_head", C.c_int),
        ("full_attn", C.c_int * FTB_MAX_STAGES),
        ("num_mem_kv", C.c_int),
        ("conditional", C.c_int),
    ]


class FtbError(RuntimeError):
    pass


if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `python __graft_entry__.py` (nvcc, sm_100a). "
        "flowtrain_stochastic_interpolation_b200 has no CPU or PyTorch fallback."
    )

lib = C.CDLL(LIB_PATH)

_vp, _i, _i64, _f, _d, _sz = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_double, C.c_size_t
_ip = C.POINTER(C.c_int)
BUCKET_CB = C.CFUNCTYPE(None, C.c_void_p, C.c_int64, C.c_int64)

_SIGS = {
    "ftb_last_error": (C.c_char_p, []),
    "ftb_version": (_i, []),
    "ftb_device_sm_count": (_i, []),
    "ftb_launch_count": (_i64, []),
    "ftb_profile_enable": (_i, [_i]),
    "ftb_profile_collect": (_i, [C.POINTER(_d), C.POINTER(_d), C.POINTER(_d), _ip, _i]),
    "ftb_unet3d_create": (_i, [C.POINTER(FtbUnetCfg), C.POINTER(_vp)]),
    "ftb_unet3d_destroy": (_i, [_vp]),
    "ftb_unet3d_num_params": (_i, [_vp]),
    "ftb_unet3d_param_name": (C.c_char_p, [_vp, _i]),
    "ftb_unet3d_param_numel": (_i64, [_vp, _i]),
    "ftb_unet3d_param_shape": (_i, [_vp, _i, _ip, _i]),
    "ftb_unet3d_set_param": (_i, [_vp, C.c_char_p, _vp, _i64, _vp]),
    "ftb_unet3d_workspace_bytes": (_sz, [_vp, _i, _i, _i, _i]),
    "ftb_unet3d_forward": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _sz, _vp]),
    "ftb_unet3d_f32_workspace_bytes": (_sz, [_vp, _i, _i, _i, _i]),
    "ftb_unet3d_forward_f32": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _sz, _vp]),
    "ftb_unet3d_cond_forward_f32": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _sz, _vp]),
    "ftb_unet3d_get_tap_f32": (_i, [_vp, C.c_char_p, _vp, _ip, _vp]),
    "ftb_unet3d_cond_workspace_bytes": (_sz, [_vp, _i, _i, _i, _i, _i]),
    "ftb_unet3d_cond_forward": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _i, _i, _i, _i, _vp, _sz, _i, _vp]),
    "ftb_unet3d_tap_channels": (_i, [_vp, C.c_char_p, _ip, _ip, _ip, _ip]),
    "ftb_unet3d_get_tap": (_i, [_vp, C.c_char_p, _vp, _vp]),
    "ftb_unet3d_last_launches": (_i, [_vp]),
    "ftb_ode_lincomb": (_i, [_vp, _vp, _vp, _vp, _i, _i64, _vp]),
    "ftb_ode_error_ratio": (_i, [_vp, _vp, _vp, _vp, _i, _f, _f, _i64, _vp, _vp]),
    "ftb_ode_scaled_sumsq": (_i, [_vp, _vp, _vp, _f, _f, _i64, _vp, _vp]),
    "ftb_ode_dense_eval": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _d, _d, _i64, _vp]),
    "ftb_ode_ctl_init": (_i, [_vp, _d, _vp]),
    "ftb_ode_ctl_first_step": (_i, [_vp, _i, _i64, _i, _vp]),
    "ftb_ode_ctl_stage_time": (_i, [_vp, _vp, _d, _i, _vp]),
    "ftb_ode_lincomb_dev": (_i, [_vp, _vp, _vp, _vp, _i, _i64, _vp, _vp]),
    "ftb_ode_error_ratio_dev": (_i, [_vp, _vp, _vp, _vp, _i, _f, _f, _i64, _vp, _vp]),
    "ftb_ode_ctl_step": (_i, [_vp, _vp, _i, _i64, _i, _i64, _vp]),
    "ftb_ode_advance": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _i, _vp, _vp, _i64, _vp]),
    "ftb_denoise_drift_dev": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i64, _vp]),
    "ftb_cond_frontend": (_i, [_vp, _vp, _vp, _i, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "ftb_cond_loss_accumulate": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i64, _vp, _vp]),
    "ftb_cond_loss_grad": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i64, _vp, _f, _f, _vp, _vp]),
    "ftb_decode_vote": (_i, [_vp, _vp, _i, _i, _i, _i64, _vp, _vp, _vp]),
    "ftb_vote_finalize": (_i, [_vp, _i, _i, _i64, _i, _vp, _vp, _vp, _vp, _vp]),
    "ftb_interp_xt_bt": (_i, [_i, _i, _f, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i64, _vp]),
    "ftb_ode_axpy": (_i, [_vp, _vp, _vp, _d, _i64, _vp, _i64, _vp]),
    "ftb_ode_heun_combine": (_i, [_vp, _vp, _vp, _vp, _d, _i64, _vp]),
    "ftb_ode_rk4_combine": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _d, _i64, _vp]),
    "ftb_denoise_drift": (_i, [_vp, _vp, _vp, _vp, _f, _f, _f, _f, _f, _i, _i64, _vp]),
    "ftb_decode": (_i, [_vp, _vp, _vp, _i, _i, _i, _i64, _vp]),
    "ftb_decode_logits": (_i, [_vp, _vp, _vp, _i, _i, _i, _i64, _vp]),
    "ftb_embed": (_i, [_vp, _vp, _vp, _i, _i, _i, _i64, _i, _vp]),
    "ftb_ema_update": (_i, [_vp, _vp, _i64, _d, _vp]),
    "ftb_mse_ratio_accumulate": (_i, [_vp, _vp, _i64, _vp, _vp]),
    "ftb_unet3d_param_offset": (_i64, [_vp, _i]),
    "ftb_unet3d_bind_params": (_i, [_vp, _vp, _vp]),
    "ftb_unet3d_mark_dirty": (_i, [_vp]),
    "ftb_unet3d_set_dropout": (_i, [_vp, _f, C.c_uint64]),
    "ftb_unet3d_train_workspace_bytes": (_sz, [_vp, _i, _i, _i, _i]),
    "ftb_unet3d_forward_train": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _sz, _vp]),
    "ftb_unet3d_cond_forward_train": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _sz, _vp]),
    "ftb_unet3d_backward": (_i, [_vp, _vp, _vp, _vp, _sz, BUCKET_CB, _vp, _vp]),
    "ftb_mse_ratio_grad": (_i, [_vp, _vp, _i64, _vp, _f, _vp, _vp]),
    "ftb_grad_sumsq": (_i, [_vp, _i64, _vp, _vp]),
    "ftb_adam_step": (_i, [_vp, _vp, _vp, _vp, _i64, _f, _f, _f, _f, _f, _i, _i, _vp, _f, _f, _vp]),
    "ftb_test_conv_wgrad": (_i, [_vp, _i, _vp, _i, _vp, _i, _i, _vp, _i, _i, _i, _i, _i, _vp]),
    "ftb_test_conv_dgrad": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _i, _i, _i, _i, _vp]),
    "ftb_test_conv3d": (_i, [_vp, _i, _vp, _i, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _i, _vp, _i, _i, _i, _i, _i, _vp]),
    "ftb_test_trilinear": (_i, [_vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "ftb_test_trilinear_bwd": (_i, [_vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
}

for _name, (_res, _args) in _SIGS.items():
    try:
        _fn = getattr(lib, _name)
    except AttributeError as e:  # pragma: no cover
        raise ImportError(f"libftb.so does not export {_name}; rebuild it") from e
    _fn.restype = _res
    _fn.argtypes = _args


def header_symbols():
    """Function names declared in include/ftb.h (used by the CPU test that checks exports)."""
    with open(HEADER_PATH) as f:
        src = f.read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ftb_[a-z0-9_]+)\s*\(", src)))


def last_error() -> str:
    return (lib.ftb_last_error() or b"").decode()


def check(rc: int):
    if rc != 0:
        raise FtbError(last_error() or f"ftb call failed with code {rc}")


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr():
    import torch

    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
