"""ctypes binding of libeventclip_b200.so (the C ABI declared in include/eventclip_b200.h).

The library is the product: there is no PyTorch-eager or CPU fallback behind these calls.  If the shared object is
missing, or the device is not a B200, importing callers get a loud error.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libeventclip_b200.so")

EC_OK = 0
EC_ERR_ARG, EC_ERR_CUDA, EC_ERR_UNSUPPORTED, EC_ERR_CAPACITY = -1, -2, -3, -4
EC_STATUS_BAD_COORD, EC_STATUS_COUNT_OVERFLOW = 1, 2
EC_FLAG_COUNT_NON_ZERO, EC_FLAG_BACKGROUND_MASK = 1, 2
EC_OUT_F32_NCHW, EC_OUT_BF16_NCHW, EC_OUT_BF16_PATCH, EC_OUT_F16_PATCH, EC_OUT_GRAY_BF16_PATCH, EC_OUT_GRAY_F16_PATCH = 0, 1, 2, 3, 4, 5
EC_EPI_BF16, EC_EPI_BF16_QGELU, EC_EPI_F32_RESADD, EC_EPI_F32, EC_EPI_PATCH, EC_EPI_F16_RESADD, EC_EPI_F16X2_RESADD = 0, 1, 2, 3, 4, 5, 6
EC_EPI_F16_OPERANDS = 0x100
EC_ATTN_CAUSAL, EC_ATTN_F16 = 1, 2
EC_AGG = {"sum": 0, "mean": 1, "max": 2}


class ECFrame(C.Structure):
    _fields_ = [("ev_start", C.c_int64), ("ev_count", C.c_int32), ("out_slot", C.c_int32)]


class ECError(RuntimeError):
    pass


_vp, _i, _i64, _f, _d = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_double
_ip = C.POINTER(C.c_int)

# name -> argtypes; every symbol include/eventclip_b200.h declares (tests/test_abi.py checks the two lists agree)
SIGNATURES = {
    "ec_last_error": ([], C.c_char_p),
    "ec_version": ([], _i),
    "ec_device_check": ([], _i),
    "ec_plan_frames": ([_vp, _i, _i64, _i, _vp, _i, _vp, _i, _vp, _vp, _ip, _ip], _i),
    "ec_event2img": ([_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp], _i),
    "ec_event2img_compact": ([_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp], _i),
    "ec_pack_events": ([_vp, _i64, _i, _i, _vp, _vp], _i),
    "ec_event2img_geometry": ([_i, _i, _ip, _ip, _ip], _i),
    "ec_center_events": ([_vp, _vp, _i, _i, _i, _vp], _i),
    "ec_flip_events": ([_vp, _vp, _vp, _i, _i, _i, _i, _vp], _i),
    "ec_gemm_bf16": ([_vp, _i, _vp, _i, _vp, _i, _i, _i, _i, _vp, _i, _vp, _i, _vp], _i),
    "ec_layernorm": ([_vp, _i64, _vp, _vp, _i, _i, _vp, _vp, _vp], _i),
    "ec_layernorm_ex": ([_vp, _i, _i64, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp], _i),
    "ec_gemm_stats_parts": ([_i], _i),
    "ec_gemm_bf16_stats": ([_vp, _i, _vp, _i, _vp, _i, _i, _i, _vp, _i, _vp, _vp, _i, _vp], _i),
    "ec_gemm_bf16_stats2": ([_vp, _i, _vp, _i, _vp, _i, _i, _i, _vp, _vp, _i, _vp, _i, _vp], _i),
    "ec_layernorm_f16x2": ([_vp, _i64, _vp, _vp, _i, _i, _vp, _vp, _vp], _i),
    "ec_gemm_ln": ([_vp, _i, _vp, _i, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _i, _vp], _i),
    "ec_row_stats_f16": ([_vp, _i64, _i, _i, _vp, _i, _vp], _i),
    "ec_attention": ([_vp, _vp, _i, _i, _i, _vp], _i),
    "ec_attention_ex": ([_vp, _vp, _i, _i, _i, _i, _vp], _i),
    "ec_embed_tokens": ([_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp], _i),
    "ec_cls_rows": ([_vp, _vp, _vp, _i, _i, _i, _vp], _i),
    "ec_f32_to_bf16": ([_vp, _vp, _i64, _vp], _i),
    "ec_im2col": ([_vp, _i, _i, _i, _i, _vp, _vp], _i),
    "ec_lora_merge": ([_vp, _vp, _vp, _i, _i, _i, _vp, _vp], _i),
    "ec_head": ([_vp, _vp, _vp, _i, _i, _i, _i, _f, _i, _i, _vp, _vp, _vp, _vp, _vp], _i),
    "ec_gemm_f32": ([_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp], _i),
    "ec_adapter_attention": ([_vp, _vp, _i, _i, _i, _i, _vp, _vp], _i),
    "ec_adapter_attention_bwd": ([_vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp], _i),
    "ec_relu_bwd": ([_vp, _vp, _vp, _i64, _vp], _i),
    "ec_blend": ([_vp, _vp, _d, _vp, _i64, _vp], _i),
    "ec_layernorm_f32": ([_vp, _vp, _vp, _i, _i, _vp, _vp], _i),
    "ec_gather_rows": ([_vp, _vp, _vp, _i, _i, _vp], _i),
    "ec_l2norm_rows": ([_vp, _vp, _i, _i, _vp], _i),
    # fine-tune step
    "ec_gemm_timing": ([_vp, _i], _i),
    "ec_gemm_splitk_choose": ([_i, _i, _i], _i),
    "ec_gemm_bf16_tn_splitk": ([_vp, _i, _vp, _i, _i, _i, _i, _i, _vp, _vp, _i, _vp], _i),
    "ec_gemm_bf16_splitk": ([_vp, _i, _vp, _i, _i, _i, _i, _i, _vp, _vp, _i, _vp], _i),
    "ec_layernorm_bwd": ([_vp, _i64, _vp, _vp, _vp, _i64, _i, _i, _vp, _i64, _vp, _vp], _i),
    "ec_quickgelu": ([_vp, _vp, _i64, _vp], _i),
    "ec_quickgelu_bwd": ([_vp, _vp, _vp, _i64, _vp], _i),
    "ec_transpose_bf16": ([_vp, _vp, _i, _i, _i64, _i64, _vp], _i),
    "ec_attention_fwd_lse": ([_vp, _vp, _vp, _i, _i, _i, _i, _vp], _i),
    "ec_attention_bwd": ([_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp], _i),
    "ec_adam": ([_vp, _vp, _vp, _vp, _i64, _f, _f, _f, _f, _f, _i, _vp], _i),
    "ec_gemm_f32_strided": ([_vp, _i64, _i64, _vp, _i64, _i64, _i, _i, _i, _f, _vp, _i64, _i, _vp], _i),
    "ec_lora_grad": ([_vp, _i64, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp], _i),
    "ec_colsum": ([_vp, _i, _i, _i, _i64, _vp, _i, _vp, _vp], _i),
    "ec_layernorm_param_grad": ([_vp, _i64, _vp, _i, _i, _vp, _i, _vp, _vp, _vp], _i),
    "ec_patch_rows_bf16": ([_vp, _i, _i, _i, _vp, _vp], _i),
    "ec_l2norm_rows_bwd": ([_vp, _vp, _vp, _i, _i, _vp, _vp], _i),
    "ec_ce_loss_bwd": ([_vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp], _i),
    "ec_probs_loss_bwd": ([_vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp], _i),
}

_lib = None


def load():
    """Loads the shared object (no CUDA call is made here)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ECError(
                f"{LIB_PATH} is missing: build it with `python -m eventclip_b200.build` "
                "(or __graft_entry__.build()). eventclip_b200 has no fallback path.")
        lib = C.CDLL(LIB_PATH)
        for name, (argtypes, restype) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.argtypes = argtypes
            fn.restype = restype
        _lib = lib
    return _lib


# number of kernel launches enqueued through the ABI (every device entry point launches exactly one kernel)
LAUNCHES = 0


def check(rc, what):
    global LAUNCHES
    if what not in ("ec_plan_frames", "ec_event2img_geometry", "ec_device_check"):
        LAUNCHES += 1
    if rc != EC_OK:
        msg = load().ec_last_error().decode("utf-8", "replace")
        raise ECError(f"{what} failed (code {rc}): {msg}")


_device_ok = set()


def require_device(device_index):
    """Raises unless the given CUDA device is sm_100 (B200).  Called once per device by the ops."""
    if device_index in _device_ok:
        return
    import torch
    if not torch.cuda.is_available():
        raise ECError("eventclip_b200 needs a CUDA device (B200, sm_100a); none is visible and there is no CPU fallback")
    with torch.cuda.device(device_index):
        check(load().ec_device_check(), "ec_device_check")
    _device_ok.add(device_index)
