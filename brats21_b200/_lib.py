"""ctypes binding of libb21.so (the C-ABI declared in include/b21.h).

There is deliberately NO fallback: if the shared library is missing or a call fails, a RuntimeError is raised.
PyTorch is only used for device memory and streams; every signature is plain pointers and integers.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import torch

_PKG = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ["B21_LIB"]) if os.environ.get("B21_LIB") else _PKG / "libb21.so"  # B21_LIB: A/B runs of two builds
STAT_SLOTS = 32

_vp, _i, _f, _d, _i64 = C.c_void_p, C.c_int, C.c_float, C.c_double, C.c_int64

# name -> argtypes (all functions return int status unless listed in _SPECIAL)
_SIGNATURES = {
    "b21_conv_cout_padded": [_i],
    "b21_pack_conv_weight": [_vp, _vp, _i, _i, _i, _i, _i, _vp],
    "b21_conv3d_fwd": [_vp, _i, _vp, _vp, _vp, _i, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp],
    "b21_pack_job_tap": [_vp, _vp, _i, _i, _i, _i, _i, _vp],
    "b21_pack_job_march": [_vp, _vp, _i, _i, _i, _vp],
    "b21_pack_job_slide": [_vp, _vp, _i, _i, _i, _vp],
    "b21_pack_batch": [_vp, _i, _i, _vp],
    "b21_conv_input_supported": [_i, _i],
    "b21_conv_input_weight_bytes": [_i],
    "b21_pack_conv_weight_input": [_vp, _vp, _i, _i, _vp],
    "b21_pack_job_input": [_vp, _vp, _i, _i, _vp],
    "b21_conv3d_input_fwd": [_vp, _i, _vp, _vp, _vp, _i, _vp, _i, _i, _i, _i, _i, _i, _vp],
    "b21_conv_march_supported": [_i, _i],
    "b21_conv_march_weight_bytes": [_i, _i],
    "b21_pack_conv_weight_march": [_vp, _vp, _i, _i, _i, _vp],
    "b21_conv3d_march_fwd": [_vp, _i, _vp, _vp, _vp, _i, _vp, _i, _i, _i, _i, _i, _i, _vp],
    "b21_conv_slide_supported": [_i, _i],
    "b21_conv_slide_weight_bytes": [_i, _i],
    "b21_pack_conv_weight_slide": [_vp, _vp, _i, _i, _i, _vp],
    "b21_conv3d_slide_fwd": [_vp, _i, _vp, _vp, _vp, _i, _vp, _i, _i, _i, _i, _i, _i, _vp],
    "b21_conv_point_supported": [_i, _i],
    "b21_conv1x1_fwd": [_vp, _i, _vp, _vp, _vp, _i, _vp, _i, _i64, _i, _i, _vp],
    "b21_norm_apply": [_vp, _i, _vp, _i, _vp, _vp, _vp, _vp, _i, _i, _i, _i64, _i, _f, _vp],
    "b21_channel_stats": [_vp, _i, _vp, _i, _i64, _i, _vp],
    "b21_norm_coeffs": [_i, _vp, _vp, _vp, _vp, _vp, _f, _vp, _vp, _i, _i, _i64, _f, _vp],
    "b21_affine_act": [_vp, _i, _vp, _i, _vp, _vp, _i, _f, _i, _i64, _i, _vp],
    "b21_se_gate": [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _f, _vp],
    "b21_scale_pool": [_vp, _i, _vp, _vp, _i, _vp, _i, _i, _i, _i, _i, _i, _i, _vp],
    "b21_upsample2x": [_vp, _i, _vp, _i, _i, _i, _i, _i, _i, _vp],
    "b21_upsample_f32": [_vp, _vp, _i, _i, _i, _i, _i, _vp],
    "b21_head_conv": [_vp, _i, _vp, _vp, _i, _vp, _vp, _vp, _i, _i64, _i, _i, _vp],
    "b21_evo_se_affine": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i64, _f, _vp],
    "b21_border_weight_sums": [_vp, _vp, _i, _i, _i, _vp],
    "b21_bias_table": [_vp, _vp, _vp, _i, _vp, _i, _i, _i, _i, _vp],
    "b21_pack_conv_weight_fold": [_vp, _vp, _i, _i, _i, _i, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp],
    "b21_pack_conv_weight_march_fold": [_vp, _vp, _i, _i, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp],
    "b21_pack_conv_weight_slide_fold": [_vp, _vp, _i, _i, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp],
    "b21_conv3d_march_fwd_fold": [_vp, _i, _vp, _i64, _vp, _vp, _vp, _i, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp],
    "b21_conv3d_march_fwd_fold2": [_vp, _i, _i, _vp, _i, _vp, _i64, _vp, _vp, _vp, _i, _vp, _vp, _i, _i, _i, _i, _i, _i, _i,
                                   _vp],
    "b21_conv3d_slide_fwd_fold": [_vp, _i, _vp, _i64, _vp, _vp, _vp, _i, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp],
    "b21_conv1x1_fwd_fold": [_vp, _i, _vp, _i, _vp, _vp, _vp, _i, _vp, _i, _i, _i64, _i, _i, _vp],
    "b21_affine_pool": [_vp, _i, _vp, _vp, _i, _vp, _i, _i, _i, _i, _i, _i, _i, _vp],
    "b21_pack_windows": [_vp, _i, _i, _i, _i, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp],
    "b21_blend_accumulate": [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _f, _vp],
    "b21_tta_accumulate": [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _i, _i, _i, _vp, _vp, _i, _i, _vp],
    "b21_labels_finalize": [_vp, _f, _f, _vp, _i, _vp, _vp, _i64, _i, _vp],
    "b21_mask_background": [_vp, _i, _vp, _i, _i64, _vp],
    "b21_conv3d_wgrad": [_vp, _i, _vp, _i, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp],
    "b21_conv_wgrad_march_supported": [_i, _i],
    "b21_conv3d_wgrad_march": [_vp, _i, _vp, _i, _vp, _i, _i, _i, _i, _i, _i, _i, _vp],
    "b21_norm_bwd_workspace_bytes": [_i, _i],
    "b21_norm_bwd": [_vp, _i, _vp, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                     _vp, _vp, _i, _vp, _i64, _i, _i, _i64, _i, _f, _vp],
    "b21_pool_bwd": [_vp, _i, _vp, _i, _vp, _i, _vp, _i, _i, _i, _i, _i, _i, _i, _vp],
    "b21_upsample2x_bwd": [_vp, _i, _vp, _i, _i, _i, _i, _i, _i, _vp],
    "b21_upsample_f32_bwd": [_vp, _vp, _i, _i, _i, _i, _i, _vp],
    "b21_head_conv_bwd": [_vp, _i, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _i, _i, _i64, _i, _i, _vp],
    "b21_add_inplace": [_vp, _i, _vp, _i, _i64, _i, _vp],
    "b21_dice_fwd": [_vp, _vp, _vp, _vp, _vp, _i, _i, _i64, _i, _f, _f, _f, _vp],
    "b21_dice_bwd": [_vp, _vp, _vp, _vp, _f, _f, _vp, _i, _i, _i64, _vp],
    "b21_ce_fwd": [_vp, _vp, _vp, _vp, _i, _i, _i64, _f, _vp],
    "b21_ranger_chunk": [],
    "b21_ranger_step": [_vp, _vp, _i, _f, _f, _f, _f, _f, _f, _f, _i, _i, _f, _vp, _vp],
    "b21_grad_centralize": [_vp, _i, _vp],
    "b21_foreground_bbox": [_vp, _i, _i, _i, _i, _vp, _vp],
    "b21_nonzero_stats": [_vp, _i, _i, _i, _i, _vp, _vp, _vp],
    "b21_normalize_crop_pad": [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _i, _i, _i, _i, _i, _i, _f, _vp],
    "b21_keep_components_workspace_bytes": [_i64],
    "b21_keep_components": [_vp, _vp, _i, _i, _i, _i, _vp],
    "b21_replace_rare_workspace_bytes": [_i],
    "b21_replace_rare_labels": [_vp, _vp, _i, _i, _i, _i, _i, _vp],
    "b21_labels_to_channels": [_vp, _vp, _i64, _vp],
}



class PackJob(C.Structure):
    """b21_pack_job of include/b21.h (64 bytes)."""
    _fields_ = [("w", C.c_void_p), ("out", C.c_void_p), ("total", C.c_longlong), ("kind", _i), ("cout", _i), ("cin", _i),
                ("tf", _i), ("p0", _i), ("p1", _i), ("p2", _i), ("p3", _i), ("blk0", _i), ("nblk", _i)]


_lib = None


def exported_symbols():
    """Every symbol include/b21.h declares (used by the CPU-side ABI test)."""
    return ["b21_last_error", "b21_version", *_SIGNATURES.keys()]


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        if os.environ.get("B21_AUTOBUILD", "1") == "1":
            from .build import build_lib
            build_lib()
        if not LIB_PATH.exists():
            raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
    lib = C.CDLL(str(LIB_PATH))
    lib.b21_last_error.restype = C.c_char_p
    lib.b21_last_error.argtypes = []
    lib.b21_version.restype = _i
    for name, argtypes in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = _i
    lib.b21_conv_march_weight_bytes.restype = C.c_longlong
    lib.b21_conv_slide_weight_bytes.restype = C.c_longlong
    lib.b21_conv_input_weight_bytes.restype = C.c_longlong
    lib.b21_keep_components_workspace_bytes.restype = C.c_longlong
    lib.b21_norm_bwd_workspace_bytes.restype = C.c_longlong
    lib.b21_replace_rare_workspace_bytes.restype = C.c_longlong
    _lib = lib
    return lib


def check(status: int, what: str):
    if status != 0:
        msg = load().b21_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"libb21 {what} failed ({status}): {msg}")


def ptr(t):
    """Raw device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


# kernel launches issued per C-ABI call (host-only helpers count 0); blend launches one kernel per window
_LAUNCHES = {"b21_conv_cout_padded": 0, "b21_conv_point_supported": 0, "b21_conv_march_supported": 0,
             "b21_conv_march_weight_bytes": 0, "b21_conv_slide_supported": 0, "b21_conv_wgrad_march_supported": 0, "b21_conv_slide_weight_bytes": 0, "b21_conv3d_fwd": 1, "b21_norm_bwd": 3, "b21_dice_fwd": 2, "b21_ce_fwd": 2,
             "b21_keep_components_workspace_bytes": 0, "b21_norm_bwd_workspace_bytes": 0, "b21_pack_job_tap": 0,
             "b21_pack_job_march": 0, "b21_pack_job_slide": 0, "b21_pack_job_input": 0, "b21_conv_input_supported": 0,
             "b21_conv_input_weight_bytes": 0, "b21_replace_rare_workspace_bytes": 0, "b21_foreground_bbox": 2,
             "b21_keep_components": 4, "b21_replace_rare_labels": 5}
launch_count = 0


def call(name: str, *args, launches: int = None):
    global launch_count
    lib = load()
    check(getattr(lib, name)(*args), name)
    launch_count += _LAUNCHES.get(name, 1) if launches is None else launches
