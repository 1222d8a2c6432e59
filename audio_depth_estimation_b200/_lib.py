"""ctypes binding of libadp_b200.so (the C ABI declared in include/adp_b200.h).

There is no CPU fallback: every numerical entry point of this package goes through this
library, and loading fails loudly when the shared object has not been built
(``python -m audio_depth_estimation_b200.build``).
"""
import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
# ADP_LIB_PATH: load another build of the same C ABI (A/B measurements against an earlier build)
LIB_PATH = os.environ.get("ADP_LIB_PATH") or os.path.join(_PKG, "libadp_b200.so")

ADP_F32 = 0
ADP_BF16 = 1
ADP_MAX_LEVELS = 10


class AdpError(RuntimeError):
    pass


class UnetDesc(C.Structure):
    _fields_ = [("batch", C.c_int), ("in_ch", C.c_int), ("out_ch", C.c_int), ("ngf", C.c_int),
                ("num_downs", C.c_int), ("size", C.c_int), ("dtype", C.c_int),
                ("final_sigmoid", C.c_int), ("training", C.c_int), ("bn_eps", C.c_float),
                ("bn_momentum", C.c_float), ("reuse_weight_cache", C.c_int), ("inference_only", C.c_int)]


_LEVEL_SLOTS = ("conv_w", "convT_w", "convT_bias", "bn_down_w", "bn_down_b", "bn_down_rm", "bn_down_rv",
                "bn_up_w", "bn_up_b", "bn_up_rm", "bn_up_rv", "conv_w_bf16", "convT_w_bf16")


class UnetLevel(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in _LEVEL_SLOTS]


class TensorRef(C.Structure):
    _fields_ = [("p", C.c_void_p), ("g", C.c_void_p), ("m", C.c_void_p), ("v", C.c_void_p), ("n", C.c_int64),
                ("p_bf16", C.c_void_p)]


_vp, _i, _f, _sz, _i64 = C.c_void_p, C.c_int, C.c_float, C.c_size_t, C.c_int64

# name -> (restype, argtypes); mirrors include/adp_b200.h one to one
SIGNATURES = {
    "adp_last_error": (C.c_char_p, []),
    "adp_version": (_i, []),
    "adp_device_is_sm100": (_i, []),
    "adp_set_tensor_core": (_i, [_i]),
    "adp_stft_mag": (_i, [_vp, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "adp_feature_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "adp_feature_forward": (_i, [_vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _sz, _vp]),
    "adp_mel_spectrogram": (_i, [_vp, _i, _i, _i, _i, _i, _i, _i, _f, _f, _f, _vp, _vp, _sz, _vp]),
    "adp_feature_mel_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "adp_feature_forward_mel": (_i, [_vp, _i, _i, _i, _i, _i, _i, _i, _f, _f, _f, _i, _i, _vp, _vp, _sz, _vp]),
    "adp_resize_aa": (_i, [_vp, _i, _i, _i, _i, _vp, _vp]),
    "adp_depth_prepare": (_i, [_vp, _i, _i, _i, _i, _i, _f, _i, _f, _vp, _vp]),
    "adp_depth_loss_sums": (_i, [_vp, _vp, _i64, _f, _f, _i, _vp, _vp]),
    "adp_depth_loss_value": (_i, [_vp, _f, _f, _f, _vp, _vp]),
    "adp_depth_loss_backward": (_i, [_vp, _vp, _i64, _f, _f, _i, _vp, _f, _f, _f, _vp, _vp, _vp]),
    "adp_depth_metrics_workspace_bytes": (_sz, [_i]),
    "adp_depth_metrics": (_i, [_vp, _vp, _i, _i64, _f, _i, _f, _f, _vp, _vp, _sz, _vp]),
    "adp_weight_operand": (_i, [_vp, _i, _i, _vp, _vp]),
    "adp_conv2d_k4s2_fprop": (_i, [_i, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "adp_conv2d_k4s2_dgrad": (_i, [_i, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "adp_conv2d_k4s2_wgrad": (_i, [_i, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "adp_convT2d_k4s2_fprop": (_i, [_i, _vp, _i, _vp, _i, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "adp_convT2d_k4s2_dgrad": (_i, [_i, _vp, _vp, _vp, _vp, _i, _vp, _i, _i, _i, _i, _i, _vp]),
    "adp_convT2d_k4s2_wgrad": (_i, [_i, _vp, _i, _vp, _i, _vp, _vp, _i, _i, _i, _i, _vp]),
    "adp_cast_bf16": (_i, [_vp, _vp, _i64, _vp]),
    "adp_conv2d_k3s1_fprop": (_i, [_vp, _i, _vp, _i, _vp, _vp, _i, _i, _i, _i, _vp, _sz, _vp]),
    "adp_conv2d_k3s1_dgrad": (_i, [_vp, _i, _vp, _vp, _i, _vp, _i, _i, _i, _i, _vp, _sz, _vp]),
    "adp_conv2d_k3s1_wgrad": (_i, [_vp, _i, _vp, _i, _vp, _i, _vp, _i, _i, _i, _vp]),
    "adp_gemm_rows_bf16": (_i, [_vp, _i, _vp, _i, _vp, _i, _vp, _i, _vp, _i, _vp, _i64, _vp]),
    "adp_conv2d_k3s1_c1_fprop": (_i, [_vp, _i64, _vp, _vp, _i, _i, _i, _i, _vp]),
    "adp_conv2d_k3s1_c1_wgrad": (_i, [_vp, _vp, _i64, _vp, _i, _i, _i, _i, _vp]),
    "adp_bn_act_forward": (_i, [_vp, _i64, _i, _vp, _vp, _vp, _vp, _vp, _i, _f, _f, _f, _vp, _vp, _vp, _vp]),
    "adp_bn_act_backward": (_i, [_vp, _i64, _i, _vp, _vp, _f, _i, _vp, _vp, _vp, _vp, _vp]),
    "adp_maxpool2_forward": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "adp_maxpool2_backward": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "adp_upsample2x_forward": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "adp_upsample2x_backward": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "adp_bilinear_resize_forward": (_i, [_vp, _vp, _i64, _i, _i, _i, _i, _vp]),
    "adp_bilinear_resize_backward": (_i, [_vp, _vp, _i64, _i, _i, _i, _i, _vp]),
    "adp_pixel_shuffle2": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "adp_rows_op": (_i, [_i, _vp, _vp, _vp, _vp, _i64, _i, _vp]),
    "adp_rows_reduce": (_i, [_i, _vp, _vp, _i64, _i, _vp, _vp, _vp]),
    "adp_softmax_rows": (_i, [_vp, _i64, _i, _f, _vp, _vp, _vp, _vp]),
    "adp_softmax_apply": (_i, [_vp, _i64, _i, _f, _vp, _vp, _i, _vp, _vp]),
    "adp_softmax_backward": (_i, [_vp, _vp, _i64, _i, _f, _vp, _i, _vp, _vp]),
    "adp_softmax_stats_init": (_i, [_vp, _vp, _i64, _vp]),
    "adp_gemm_rows_softmax": (_i, [_vp, _i, _vp, _i, _i64, _i, _i, _f, _vp, _vp, _vp, _vp, _vp, _vp]),
    "adp_gemm_tn_bf16": (_i, [_vp, _i, _vp, _i, _vp, _i, _i64, _vp]),
    "adp_depth_head_forward": (_i, [_vp, _vp, _vp, _f, _i64, _i, _vp, _vp]),
    "adp_depth_head_backward": (_i, [_vp, _vp, _vp, _f, _vp, _i64, _i, _vp, _vp, _vp, _vp]),
    "adp_first_conv_k4s2_fprop": (_i, [_vp, _vp, _vp, _vp, _f, _vp, _f, _i, _i, _i, _vp]),
    "adp_first_conv_k4s2_wgrad": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp]),
    "adp_first_conv_k4s2_wgrad_act": (_i, [_vp, _vp, _vp, _vp, _f, _vp, _i, _i, _i, _vp]),
    "adp_last_convT_k4s2_dgrad": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "adp_last_convT_k4s2_wgrad": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "adp_last_convT_k4s2_fprop": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _i, _i, _i, _vp]),
    "adp_set_option": (_i, [C.c_char_p, _i]),
    "adp_profile_read_n": (_i, [_i, _vp, _vp, _vp]),
    "adp_selftest_umma_offset": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp]),
    "adp_unet_workspace_bytes": (_sz, [C.POINTER(UnetDesc)]),
    "adp_unet_forward": (_i, [C.POINTER(UnetDesc), _vp, C.POINTER(UnetLevel), _vp, _sz, _vp, _vp]),
    "adp_unet_backward": (_i, [C.POINTER(UnetDesc), _vp, _vp, _vp, C.POINTER(UnetLevel), C.POINTER(UnetLevel),
                               _vp, _sz, _vp]),
    "adp_unet_backward_stages": (_i, [C.POINTER(UnetDesc), _vp, _vp, _vp, C.POINTER(UnetLevel),
                                      C.POINTER(UnetLevel), _vp, _sz, _i, _i, _vp]),
    "adp_launch_count": (C.c_longlong, []),
    "adp_tc_launch_count": (C.c_longlong, []),
    "adp_profile_enable": (_i, [_i]),
    "adp_profile_read": (_i, [_vp, _vp, _vp]),
    "adp_grad_sumsq": (_i, [C.POINTER(TensorRef), _i, _vp, _vp]),
    "adp_clip_adamw_step": (_i, [C.POINTER(TensorRef), _i, _vp, _f, _f, _f, _f, _f, _f, _i, _vp, _vp]),
    "adp_clip_adamw_step_graph": (_i, [C.POINTER(TensorRef), _i, _vp, _f, _f, _f, _f, _f, _f, _vp, _vp, _vp, _vp]),
}

_lib = None


def load():
    """Load the shared library (once) and attach the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise AdpError("libadp_b200.so is not built: run `python -m audio_depth_estimation_b200.build` "
                       "(this package has no CPU or PyTorch fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise AdpError("libadp_b200 error %d: %s" % (rc, load().adp_last_error().decode()))


def stream_ptr():
    import torch
    return torch.cuda.current_stream().cuda_stream


def on_device(t):
    """Context manager: make `t`'s GPU the current device for the library call (launches, per-device kernel attributes and
    `stream_ptr()` all refer to the CURRENT device; a process may drive several GPUs, e.g. gpu_ids=[1])."""
    import torch
    return torch.cuda.device(t.device)


def require_cuda(t, name="tensor", dtype=None):
    """The product path is CUDA only: refuse CPU tensors instead of falling back."""
    if not t.is_cuda:
        raise AdpError("%s must be a CUDA tensor: audio_depth_estimation_b200 has no CPU path" % name)
    if dtype is not None and t.dtype != dtype:
        raise AdpError("%s must have dtype %s, got %s" % (name, dtype, t.dtype))
    return t
