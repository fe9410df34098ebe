"""ctypes binding of libcuriosity_b200.so (the C ABI declared in include/curiosity_b200.h).

There is no CPU fallback: importing this module without the built library, or calling a
compute entry point without a CUDA device, raises.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcuriosity_b200.so")

U8, BF16, F32 = 0, 1, 2
ACT_NONE, ACT_RELU, ACT_LRELU, ACT_ELU = 0, 1, 2, 3
ENGINE_AUTO, ENGINE_SIMT, ENGINE_UMMA = 0, 1, 2


class ConvDesc(C.Structure):
    _fields_ = [("n", C.c_int32), ("in_h", C.c_int32), ("in_w", C.c_int32), ("in_c", C.c_int32),
                ("k_h", C.c_int32), ("k_w", C.c_int32), ("stride", C.c_int32), ("pad", C.c_int32),
                ("out_h", C.c_int32), ("out_w", C.c_int32), ("out_c", C.c_int32),
                ("in_dtype", C.c_int32),
                ("in_stride_n", C.c_int64), ("in_stride_h", C.c_int64),
                ("in_stride_w", C.c_int64), ("in_stride_c", C.c_int64),
                ("act", C.c_int32), ("engine", C.c_int32)]


class Epilogue(C.Structure):
    _fields_ = [("bias", C.c_void_p), ("residual", C.c_void_p), ("side", C.c_void_p),
                ("side_w", C.c_void_p), ("side_dim", C.c_int32), ("norm_mean", C.c_void_p),
                ("norm_rstd", C.c_void_p), ("out_bf16", C.c_void_p), ("out_f32", C.c_void_p),
                ("side_bf16", C.c_void_p)]


class CuriosityLibError(RuntimeError):
    pass


_P, _I, _L, _F = C.c_void_p, C.c_int32, C.c_int64, C.c_float
_SIGS = {
    "cb_discount_return": [_P, _P, _P, _F, _P, _I, _I, _P],
    "cb_gae": [_P, _P, _P, _P, _F, _F, _P, _P, _I, _I, _P],
    "cb_valid_from_done": [_P, _P, _I, _I, _P],
    "cb_valid_lengths": [_P, _P, _P, _I, _I, _P],
    "cb_obs_moments_u8": [_P, _L, _I, _P, _I, _I, _P, _P, _P],
    "cb_rms_merge_from_sums": [_P, _P, _P, _P, _P, _P, _P, _P, _I, _P],
    "cb_frames_differ": [_P, _L, _P, _L, _I, _L, _P, _P],
    "cb_reward_filter": [_P, _P, _F, _P, _P, _P, _I, _I, _P],
    "cb_moments_partial": [_P, _L, _P, _P, _P],
    "cb_rms_merge_moments": [_P, _P, C.c_double, _P, _P, _P, _P],
    "cb_valid_count": [_P, _P, _L, _P],
    "cb_masked_normalize": [_P, _P, _F, _P, _L, _P],
    "cb_rms_update_columns": [_P, _L, _I, _P, _P, _P, _P],
    "cb_masked_moments": [_P, _P, _L, _P, _P, _P],
    "cb_normalize_apply": [_P, _P, _F, _P, _L, _P],
    "cb_rnd_scale_bonus": [_P, _P, _P, _F, _P, _L, _P],
    "cb_conv_fwd": [C.POINTER(ConvDesc), _P, _P, C.POINTER(Epilogue), _P],
    "cb_conv_dgrad": [C.POINTER(ConvDesc), _P, _P, _P, _I, _P, _P, _P],
    "cb_conv_dgrad_s2_supported": [C.POINTER(ConvDesc)],
    "cb_conv_dgrad_s2": [C.POINTER(ConvDesc), _P, _P, _P, _I, _P, _P, _P],
    "cb_conv1_supported": [C.POINTER(ConvDesc)],
    "cb_conv1_fwd": [C.POINTER(ConvDesc), _P, _P, C.POINTER(Epilogue), _P, _P, _P, _P],
    "cb_rnd_conv12_supported": [],
    "cb_bn_moments": [_P, _L, _I, _P, _P, _P],
    "cb_bn_apply": [_P, _L, _I, _P, _P, _P, _P, _P],
    "cb_bn_bwd_moments": [_P, _L, _I, _P, _P, _P, _P, _P],
    "cb_bn_bwd_apply": [_P, _L, _I, _P, _P, _P, _P, _I, _P, _P],
    "cb_rnd_normalize_s2d": [_P, _L, _I, _P, _P, _P, _P],
    "cb_rnd_conv12_fwd": [_P, _I, _P, _P, _P, _P, _P, _P, _P],
    "cb_rnd_conv12_twin_fwd": [_P, _I] + [_P] * 12,
    "cb_conv1_wgrad": [C.POINTER(ConvDesc), _P, _P, _P, _P, _P, _P, _F, _P],
    "cb_conv_wgrad": [C.POINTER(ConvDesc), _P, _P, _P, _P, _P, _P, _F, _P],
    "cb_side_wgrad": [_P, _P, _P, _L, _I, _I, _F, _P],
    "cb_im2col": [C.POINTER(ConvDesc), _P, _P, _I, _I, _I, _P],
    "cb_col2im": [C.POINTER(ConvDesc), _P, _I, _P, _I, _P, _P, _L, _L, _L, _I, _I, _P],
    "cb_colsum_bf16": [_P, _L, _I, _P, _F, _P],
    "cb_grouped_linear_fwd": [_I, _I, _I, _I, _I, _P, _P, _I, C.POINTER(Epilogue), _P],
    "cb_grouped_linear_dgrad": [_I, _I, _I, _I, _P, _P, _P, _I, _P, _P, _P],
    "cb_grouped_linear_wgrad": [_I, _I, _I, _I, _I, _P, _P, _P, _P, _F, _P],
    "cb_ensemble_sqerr": [_P, _P, _P, _F, _F, _P, _L, _I, _I, _P],
    "cb_ensemble_sqerr_grad": [_P, _P, _P, _F, _P, _F, _P, _L, _I, _I, _P],
    "cb_ensemble_variance_packed": [_P, _I, _F, _P, _L, _I, _P],
    "cb_rowwise_sqerr": [_P, _P, _F, _P, _L, _I, _P],
    "cb_sqerr_grad": [_P, _P, _P, _F, _P, _F, _P, _L, _I, _P],
    "cb_rowwise_sqerr_masked": [_P, _P, _P, _F, _F, _P, _L, _I, _P],
    "cb_ensemble_variance": [_P, _I, _F, _P, _L, _I, _P],
    "cb_cross_entropy": [_P, _P, _P, _P, _P, _L, _I, _P],
    "cb_bce_logits": [_P, _P, _F, _P, _F, _P, _L, _I, _P],
    "cb_bce_horizons": [_P, _P, _I, _I, _I, _I, _I, _I, C.POINTER(C.c_int32), _P, _F, _P, _P],
    "cb_action_windows": [_P, _I, _I, _I, _I, _P, _I, _P, _P],
    "cb_weighted_mean": [_P, _P, _P, _L, _P],
    "cb_loss_coef": [_P, _P, _F, _P, _P, _L, _P],
    "cb_dot": [_P, _P, _P, _L, _P],
    "cb_lstm_cell_fwd": [_P, _P, _P, _P, _P, _P, _P, _I, _I, _P],
    "cb_lstm_cell_bwd": [_P, _P, _P, _P, _P, _P, _P, _I, _I, _P],
    "cb_lstm_sequence_fwd": [_I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _I, _P],
    "cb_lstm_sequence_bwd": [_I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _I, _P],
    "cb_gru_cell_fwd": [_P, _P, _P, _P, _P, _P, _P, _I, _I, _P],
    "cb_gru_cell_bwd": [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _P],
    "cb_gru_sequence_fwd": [_I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _I, _P],
    "cb_gru_sequence_bwd": [_I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _I, _P],
    "cb_gru_persistent_supported": [_I],
    "cb_gru_persistent_fwd": [_I, _I, _I, _P, _P, _P, _P, _P, _P, _P],
    "cb_gru_persistent_bwd": [_I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P],
    "cb_softmax_fwd": [_P, _P, _L, _I, _P],
    "cb_softmax_bwd": [_P, _P, _P, _L, _I, _P],
    "cb_ppo_loss": [_P, _P, _P, _P, _P, _P, _P, _F, _F, _F, _P, _P, _P, _P, _L, _I, _P],
    "cb_add_to_bf16": [_P, _P, _P, _L, _P],
    "cb_policy_heads_supported": [_I, _I],
    "cb_policy_heads_fwd": [_P, _P, _P, _P, _P, _P, _P, _P, _L, _I, _I, _P],
    "cb_policy_heads_bwd": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _L, _I, _I, _I, _P],
    "cb_sumsq": [_P, _L, _P, _P],
    "cb_adam_clip_step": [_P, _P, _P, _P, _L, _P, _F, _F, _F, _F, _F, _I, _P],
    "cb_adam_clip_step_dev": [_P, _P, _P, _P, _L, _P, _F, _F, _F, _F, _P, _P, _P],
    "cb_h2d_2d": [_P, _L, _P, _L, _L, _L, _P],
    "cb_prepare_conv_weight": [_P, _P, _P, _I, _I, _I, _I, _P],
    "cb_unprepare_conv_grad": [_P, _P, _I, _I, _I, _I, _F, _P],
}
EXPORTS = sorted(list(_SIGS) + ["cb_last_error", "cb_version", "cb_device_supports_umma"])

_lib = None


def load():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CuriosityLibError(
            f"{LIB_PATH} is missing: build it with `python -m curiosity_baselines_b200.build` "
            "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, args in _SIGS.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = C.c_int
    lib.cb_last_error.restype = C.c_char_p
    lib.cb_version.restype = C.c_int
    lib.cb_device_supports_umma.restype = C.c_int
    _lib = lib
    return lib


def ptr(t):
    """Device pointer of a tensor (None -> NULL).  Tensors must be CUDA and contiguous."""
    if t is None:
        return None
    if not t.is_cuda:
        raise CuriosityLibError("expected a CUDA tensor (no CPU fallback)")
    if not t.is_contiguous():
        raise CuriosityLibError("expected a contiguous tensor")
    return t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


_launches = 0


def launch_count():
    """Number of library entry-point calls made so far (each launches >= 1 of our kernels)."""
    return _launches


def call(name, *args):
    global _launches
    lib = load()
    rc = getattr(lib, name)(*args)
    _launches += 1
    if rc != 0:
        raise CuriosityLibError(f"{name} failed ({rc}): {lib.cb_last_error().decode()}")


_probe = None


def load_probe():
    """The kernel author's descriptor probe (csrc/umma_probe.cu), a separate test-only library."""
    global _probe
    if _probe is None:
        path = os.path.join(_HERE, "libcuriosity_b200_probe.so")
        if not os.path.exists(path):
            raise CuriosityLibError(f"{path} is missing: `python -m curiosity_baselines_b200.build --probe`")
        lib = C.CDLL(path)
        lib.cb_umma_probe.argtypes = [_P, _I, _P, _I, C.POINTER(C.c_int), C.POINTER(C.c_int), _I, _I, _I, _I, _I, _P, _P]
        lib.cb_umma_probe.restype = C.c_int
        lib.cb_probe_last_error.restype = C.c_char_p
        _probe = lib
    return _probe


def call_probe(*args):
    lib = load_probe()
    rc = lib.cb_umma_probe(*args)
    if rc != 0:
        raise CuriosityLibError(f"cb_umma_probe failed ({rc}): {lib.cb_probe_last_error().decode()}")
