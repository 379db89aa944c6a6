"""ctypes binding of libvr2.so (include/vr2.h).  Loading fails loudly: there is no
Python / PyTorch fallback for any entry point."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VR2_LIB") or os.path.join(_HERE, "libvr2.so")   # VR2_LIB: an alternative build (A/B timing)

VR2_F32, VR2_BF16, VR2_I64, VR2_I32, VR2_U8, VR2_F16 = 0, 1, 2, 3, 4, 5
VR2_PRECISION_BF16, VR2_PRECISION_SPLIT, VR2_PRECISION_F16 = 0, 1, 2
PRECISIONS = {"bf16": VR2_PRECISION_BF16, "split": VR2_PRECISION_SPLIT, "f16": VR2_PRECISION_F16}
# host-side policies on top of the library's three precisions: "auto" = f16 with the saturation guard read after every
# forward and an automatic re-run in split when anything was clamped (the default: meets max-abs 2e-2 at full speed on
# in-range activations, and is never silently wrong on outliers)
HOST_PRECISIONS = dict(PRECISIONS, auto=VR2_PRECISION_F16)


class Vr2Error(RuntimeError):
    pass


class PlanConfig(C.Structure):
    _fields_ = [
        ("spatial_merge_size", C.c_int32),
        ("window_size", C.c_int32),
        ("patch_size", C.c_int32),
        ("head_dim", C.c_int32),
        ("rope_theta", C.c_float),
    ]


class TowerConfig(C.Structure):
    _fields_ = [
        ("depth", C.c_int32),
        ("hidden_size", C.c_int32),
        ("num_heads", C.c_int32),
        ("intermediate_size", C.c_int32),
        ("out_hidden_size", C.c_int32),
        ("patch_size", C.c_int32),
        ("temporal_patch_size", C.c_int32),
        ("in_channels", C.c_int32),
        ("spatial_merge_size", C.c_int32),
        ("window_size", C.c_int32),
        ("n_fullatt", C.c_int32),
        ("fullatt_block_indexes", C.c_int32 * 8),
        ("rms_norm_eps", C.c_float),
        ("rope_theta", C.c_float),
    ]


class RopeConfig(C.Structure):
    _fields_ = [
        ("spatial_merge_size", C.c_int32),
        ("image_token_id", C.c_int64),
        ("video_token_id", C.c_int64),
        ("vision_start_token_id", C.c_int64),
        ("tokens_per_second", C.c_float),
        ("spg_int_cast", C.c_int32),
    ]


# name -> (restype, argtypes); mirrors include/vr2.h one to one
_P = C.c_void_p
SIGNATURES = {
    "vr2_last_error": (C.c_char_p, []),
    "vr2_abi_version": (C.c_int, []),
    "vr2_plan_create": (C.c_int, [C.POINTER(PlanConfig), C.POINTER(C.c_int64), C.c_int32, _P, C.POINTER(_P)]),
    "vr2_plan_destroy": (None, [_P]),
    "vr2_plan_num_patches": (C.c_int64, [_P]),
    "vr2_plan_num_tokens": (C.c_int64, [_P]),
    "vr2_plan_num_windows": (C.c_int64, [_P]),
    "vr2_plan_num_slices": (C.c_int64, [_P]),
    "vr2_plan_export": (C.c_int, [_P, _P, _P, _P, _P, _P, _P]),
    "vr2_plan_export_rotary": (C.c_int, [_P, _P, _P, _P]),
    "vr2_gather_rows": (C.c_int, [_P, _P, _P, C.c_int64, C.c_int64, C.c_int64, _P]),
    "vr2_tower_create": (C.c_int, [C.POINTER(TowerConfig), C.c_int32, C.POINTER(_P)]),
    "vr2_tower_destroy": (None, [_P]),
    "vr2_tower_set_weight": (C.c_int, [_P, C.c_char_p, _P, C.c_int32, C.POINTER(C.c_int64), C.c_int32, _P]),
    "vr2_tower_finalize": (C.c_int, [_P]),
    "vr2_tower_set_precision": (C.c_int, [_P, C.c_int32]),
    "vr2_tower_get_precision": (C.c_int32, [_P]),
    "vr2_tower_workspace_bytes": (C.c_size_t, [_P, C.c_int64]),
    "vr2_tower_forward": (C.c_int, [_P, _P, _P, C.c_int32, _P, C.c_int32, _P, C.c_size_t, _P]),
    "vr2_tower_forward_multi": (C.c_int, [_P, _P, _P, C.c_int32, C.POINTER(C.c_void_p), C.c_int32, _P, C.c_size_t, _P]),
    "vr2_tower_forward_scatter": (C.c_int, [_P, _P, _P, C.c_int32, _P, C.c_int32, C.c_int64, _P, _P, C.c_size_t, _P]),
    "vr2_token_positions_scratch_bytes": (C.c_size_t, [C.c_int64]),
    "vr2_token_positions": (C.c_int, [_P, C.c_int64, C.c_int64, _P, C.c_int64, _P, C.POINTER(C.c_int64), _P]),
    "vr2_token_positions_count": (C.c_int, [_P, C.c_int64, C.POINTER(C.c_int64), _P]),
    "vr2_scatter_rows": (C.c_int, [_P, C.c_int32, _P, C.c_int32, _P, C.c_int64, C.c_int64, C.c_int64, _P]),
    "vr2_patchify": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                               C.POINTER(C.c_float), C.POINTER(C.c_float), _P, C.c_int32, _P]),
    "vr2_resize_workspace_bytes": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_size_t)]),
    "vr2_resize_frames": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_int64, C.c_int64, C.c_int64,
                                    C.c_int32, C.c_int32, _P, C.c_int32, _P, _P]),
    "vr2_mrope_apply": (C.c_int, [_P, _P, _P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P, _P,
                                  C.POINTER(C.c_int32), C.c_float, _P]),
    "vr2_mrope_table": (C.c_int, [_P, C.c_int32, C.c_int32, _P, C.c_int32, C.POINTER(C.c_int32), C.c_float, _P, _P, _P]),
    "vr2_decoder_qkv_mrope": (C.c_int, [_P, _P, _P, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P, _P, _P, _P]),
    "vr2_lm_logprob_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int32]),
    "vr2_lm_logprob": (C.c_int, [_P, _P, _P, C.c_int64, C.c_int32, C.c_int32, C.c_float, _P, _P, C.c_size_t, _P]),
    "vr2_tower_saturation_count": (C.c_int, [_P, C.POINTER(C.c_uint64), C.c_int32, _P]),
    "vr2_tower_saturation_snapshot": (C.c_int, [_P, _P, _P]),
    "vr2_tower_saturation_counts": (C.c_int, [_P, C.POINTER(C.c_uint64), C.c_int32, _P]),
    "vr2_tower_set_f16_shift": (C.c_int, [_P, C.c_int32, C.c_int32]),
    "vr2_tower_get_f16_shift": (C.c_int, [_P, C.POINTER(C.c_int32)]),
    "vr2_tower_set_debug_stop": (C.c_int, [_P, C.c_int32]),
    "vr2_tower_workspace_layout": (C.c_int, [_P, C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "vr2_tower_profile": (C.c_int, [_P, C.c_uint32]),
    "vr2_tower_profile_read": (C.c_int, [_P, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "vr2_tower_launch_count": (C.c_int64, [_P, _P, C.c_int32]),
    "vr2_rope_scratch_bytes": (C.c_size_t, [C.c_int32, C.c_int32]),
    "vr2_rope_index": (C.c_int, [C.POINTER(RopeConfig), _P, _P, C.c_int32, C.c_int32, C.c_int32, _P, C.c_int32, _P,
                                 C.c_int32, _P, _P, _P, _P, _P]),
    "vr2_rope_index_status": (C.c_int, [_P, C.c_int32, C.c_int32, C.POINTER(C.c_int32), _P]),
    "vr2_op_linear": (C.c_int, [_P, _P, _P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P]),
    "vr2_op_linear_f16": (C.c_int, [_P, _P, _P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P]),
    "vr2_op_linear_split": (C.c_int, [_P, _P, _P, _P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P]),
    "vr2_op_rmsnorm": (C.c_int, [_P, _P, _P, C.c_int64, C.c_int32, C.c_float, _P]),
    "vr2_op_attention": (C.c_int, [_P, _P, C.POINTER(C.c_int32), C.c_int32, C.c_int32, _P]),
    "vr2_op_attention_f16": (C.c_int, [_P, _P, C.POINTER(C.c_int32), C.c_int32, C.c_int32, _P]),
    "vr2_op_attention_split": (C.c_int, [_P, _P, _P, _P, C.POINTER(C.c_int32), C.c_int32, C.c_int32, _P]),
}

_lib = None


def lib():
    """Loads libvr2.so once.  Raises if it has not been built (`python -c 'import
    __graft_entry__ as g; g.build()'` or `make -C video_r2_b200/csrc`)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise Vr2Error(
                f"{LIB_PATH} is missing: build it with `make -C video_r2_b200/csrc` "
                "(nvcc, sm_100a). video_r2_b200 has no CPU or PyTorch fallback."
            )
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the .so is stale
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc: int, what: str = ""):
    if rc == 0:
        return
    msg = lib().vr2_last_error().decode("utf-8", "replace")
    if rc == -1:
        raise ValueError(f"{what}: {msg}" if what else msg)
    raise Vr2Error(f"{what}: {msg} (status {rc})" if what else f"{msg} (status {rc})")
