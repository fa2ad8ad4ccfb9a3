"""ctypes binding of libctn_b200.so (include/ctn_b200.h).  No CPU fallback: if the library is missing or a
tensor is not on a CUDA device the call raises."""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# CTN_B200_LIB points at an alternative build of the library (A/B of compile-time tuning constants); default in-tree
_LIB_PATH = os.environ.get("CTN_B200_LIB") or os.path.join(_HERE, "libctn_b200.so")

c_i32, c_i64, c_f32, c_vp = ctypes.c_int32, ctypes.c_int64, ctypes.c_float, ctypes.c_void_p


class CtnConfig(ctypes.Structure):
    _fields_ = [(k, c_i32) for k in ("N", "L", "B", "H", "P", "X", "R", "C", "norm_type", "causal", "mask_nonlinear")]


_P = ctypes.POINTER(CtnConfig)

# name -> (restype, argtypes); must list every symbol include/ctn_b200.h declares (tests check it)
SIGNATURES = {
    "ctn_version": (c_i32, []),
    "ctn_last_error": (ctypes.c_char_p, []),
    "ctn_launch_count": (c_i64, []),
    "ctn_timing_report": (c_i32, [c_i32]),
    "ctn_param_tensors": (c_i32, [_P]),
    "ctn_param_floats": (c_i64, [_P]),
    "ctn_param_layout": (c_i32, [_P, ctypes.POINTER(c_i64), ctypes.POINTER(c_i64), c_i32]),
    "ctn_num_frames": (c_i32, [_P, c_i32]),
    "ctn_workspace_bytes": (c_i64, [_P, c_i32, c_i32, c_i32]),
    "ctn_grad_bucket": (c_i32, [_P, c_i32, ctypes.POINTER(c_i64), ctypes.POINTER(c_i64)]),
    "ctn_model_forward": (c_i32, [_P, c_vp, c_vp, c_i32, c_i32, c_vp, c_vp, c_i64, c_i32, c_vp]),
    "ctn_norm_state_floats": (c_i64, [_P]),
    "ctn_assemble_batch": (c_i32, [c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp]),
    "ctn_pack_valid": (c_i32, [c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_vp, c_vp]),
    "ctn_model_forward_bn": (c_i32, [_P, c_vp, c_vp, c_vp, c_i32, c_i32, c_vp, c_vp, c_i64, c_i32, c_i32, c_vp]),
    "ctn_model_backward": (c_i32, [_P, c_vp, c_vp, c_i32, c_i32, c_vp, c_vp, c_vp, c_i64, c_i32, c_vp]),
    "ctn_model_backward_stage": (c_i32, [_P, c_vp, c_vp, c_i32, c_i32, c_vp, c_vp, c_vp, c_i64, c_i32, c_i32, c_vp]),
    "ctn_pit_workspace_bytes": (c_i64, [c_i32, c_i32]),
    "ctn_pit_forward": (c_i32, [c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "ctn_pit_backward": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_vp, c_vp]),
    "ctn_sisnri_workspace_bytes": (c_i64, [c_i32, c_i32]),
    "ctn_sisnri": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp]),
    "ctn_reorder_source": (c_i32, [c_vp, c_vp, c_i32, c_i32, c_i64, c_vp, c_vp]),
    "ctn_overlap_and_add": (c_i32, [c_vp, c_i64, c_i32, c_i32, c_i32, c_vp, c_vp]),
    "ctn_peer_alloc": (c_i32, [c_i64, ctypes.POINTER(c_vp)]),
    "ctn_peer_free": (c_i32, [c_vp]),
    "ctn_peer_export": (c_i32, [c_vp, ctypes.c_char_p]),
    "ctn_peer_open": (c_i32, [ctypes.c_char_p, ctypes.POINTER(c_vp)]),
    "ctn_peer_close": (c_i32, [c_vp]),
    "ctn_peer_all_reduce": (c_i32, [ctypes.POINTER(c_vp), ctypes.POINTER(c_vp), c_i32, c_i32, c_i64, c_i64, c_f32, c_vp]),
    "ctn_clip_grad_norm": (c_i32, [c_vp, c_i64, c_f32, c_vp, c_vp, c_vp]),
    "ctn_adam_step": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_i64, c_f32, c_f32, c_f32, c_f32, c_f32, c_vp, c_vp]),
    "ctn_encoder_fwd": (c_i32, [c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp]),
    "ctn_encoder_bwd": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp]),
    "ctn_row_stats": (c_i32, [c_vp, c_vp, c_i64, c_i32, c_vp, c_vp]),
    "ctn_conv1x1": (c_i32, [c_vp, c_vp, c_i32, c_vp, c_i64, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp,
                            c_vp, c_vp, c_vp]),
    "ctn_conv1x1_planes": (c_i32, [c_vp, c_vp, c_vp, c_i32, c_vp, c_i64, c_i32, c_i32, c_i32, c_vp]),
    "ctn_wgrad": (c_i32, [c_vp, c_vp, c_vp, c_i64, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "ctn_prep_normfold": (c_i32, [c_vp, c_vp, c_vp, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp]),
    "ctn_dwconv_fwd": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32,
                               c_vp, c_vp, c_vp, c_vp]),
    "ctn_dwconv_bwd": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32,
                               c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "ctn_dwconv_bwd_gln_fused": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i32,
                                         c_i32, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "ctn_norm_bwd_reduce": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp]),
    "ctn_norm_bwd_apply": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_vp, c_vp]),
    "ctn_batchnorm_stats": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp,
                                    c_vp]),
    "ctn_batchnorm_bwd": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i32, c_i64, c_i32, c_vp, c_vp, c_vp,
                                  c_vp, c_vp]),
    "ctn_decoder_fwd": (c_i32, [c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp]),
    "ctn_decoder_bwd": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp,
                                c_vp, c_vp]),
}

_lib = None


def lib():
    """The loaded library; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise RuntimeError(
                f"{_LIB_PATH} is missing: the CUDA extension is not built (run `python -m conv_tasnet_b200.build`). "
                "conv_tasnet_b200 has no CPU or PyTorch fallback.")
        L = ctypes.CDLL(_LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise RuntimeError("libctn_b200: " + lib().ctn_last_error().decode())


def ptr(t):
    """device pointer of a contiguous CUDA tensor (None -> NULL)"""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("conv_tasnet_b200 runs on CUDA tensors only (there is no CPU path); got a CPU tensor")
    if not t.is_contiguous():
        raise RuntimeError("internal error: non-contiguous tensor passed to the C ABI")
    return t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def make_config(N, L, B, H, P, X, R, C, norm_type, causal, mask_nonlinear):
    norm = {"gLN": 0, "cLN": 1}.get(norm_type, 2)  # anything else is BatchNorm, like chose_norm (src/conv_tasnet.py:306)
    mask = {"relu": 0, "softmax": 1}.get(mask_nonlinear, -1)
    return CtnConfig(N, L, B, H, P, X, R, C, norm, 1 if causal else 0, mask)
