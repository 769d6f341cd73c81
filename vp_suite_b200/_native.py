"""
ctypes binding of libvpk.so (the C ABI declared in include/vpk.h).

The library is built in-tree by ``vp_suite_b200/csrc/Makefile`` (``python -c "import __graft_entry__ as g; g.build()"``).
There is no fallback: if the shared object is missing, or a compute entry is called without a CUDA device,
the call raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# VPK_LIB_PATH: a developer build of the library (e.g. the -DVPK_TRACE phase-timeline build); bench.py records the variable
LIB_PATH = os.environ.get("VPK_LIB_PATH") or os.path.join(_HERE, "libvpk.so")

VPK_PREC_FP32, VPK_PREC_BF16 = 0, 1
VPK_BACKEND_AUTO, VPK_BACKEND_SIMT = 0, 1
(VPK_MODEL_CONVLSTM_SHI, VPK_MODEL_PREDRNN_PP, VPK_MODEL_PHY, VPK_MODEL_CONVLSTM_BRANCH, VPK_MODEL_ST_PHY,
 VPK_MODEL_TRAJGRU, VPK_MODEL_PREDRNN_PP_CAUSAL) = 0, 1, 2, 3, 4, 5, 6

PRECISIONS = {"fp32": VPK_PREC_FP32, "bf16": VPK_PREC_BF16}
BACKENDS = {"auto": VPK_BACKEND_AUTO, "simt": VPK_BACKEND_SIMT}


class ModelDesc(C.Structure):
    """Mirror of ``vpk_model_desc`` (include/vpk.h); field order and types must match exactly."""
    _fields_ = [
        ("kind", C.c_int32), ("precision", C.c_int32), ("backend", C.c_int32),
        ("img_c", C.c_int32), ("img_h", C.c_int32), ("img_w", C.c_int32),
        ("enc_c", C.c_int32 * 6), ("dec_c", C.c_int32 * 6),
        ("enc_conv_k", C.c_int32 * 3), ("enc_conv_s", C.c_int32 * 3), ("enc_conv_p", C.c_int32 * 3),
        ("dec_conv_k", C.c_int32 * 3), ("dec_conv_s", C.c_int32 * 3), ("dec_conv_p", C.c_int32 * 3),
        ("enc_rnn_k", C.c_int32 * 3), ("dec_rnn_k", C.c_int32 * 3),
        ("final_conv_c", C.c_int32), ("ef_act", C.c_int32),
        ("patch_size", C.c_int32), ("num_layers", C.c_int32), ("num_hidden", C.c_int32 * 8),
        ("filter_size", C.c_int32), ("decoupling_loss_scale", C.c_float), ("layer_norm", C.c_int32),
        ("phycell_n_layers", C.c_int32), ("phycell_channels", C.c_int32), ("phycell_kernel_size", C.c_int32),
        ("convlstm_n_layers", C.c_int32), ("convlstm_hidden_dims", C.c_int32 * 8),
        ("convlstm_kernel_size", C.c_int32),
        ("max_microbatch", C.c_int32), ("use_cuda_graph", C.c_int32),
        ("action_conditional", C.c_int32), ("action_size", C.c_int32), ("residual_on_action_conv", C.c_int32),
        ("enc_rnn_L", C.c_int32 * 3), ("dec_rnn_L", C.c_int32 * 3),
        ("inflated_action_dim", C.c_int32),
    ]


# every symbol include/vpk.h declares: name -> (restype, argtypes)
_vp = C.c_void_p
_fp = C.POINTER(C.c_float)
SYMBOLS = {
    "vpk_model_create": (C.c_int, [C.POINTER(ModelDesc), C.POINTER(_vp)]),
    "vpk_model_set_param": (C.c_int, [_vp, C.c_char_p, _vp, C.POINTER(C.c_int64), C.c_int32]),
    "vpk_model_num_params": (C.c_int, [_vp, C.POINTER(C.c_int32)]),
    "vpk_model_param_info": (C.c_int, [_vp, C.c_int32, C.POINTER(C.c_char_p), C.POINTER(C.c_int64),
                                       C.POINTER(C.c_int32)]),
    "vpk_model_finalize": (C.c_int, [_vp, _vp]),
    "vpk_model_workspace_bytes": (C.c_int, [_vp, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_size_t)]),
    "vpk_model_forward": (C.c_int, [_vp, _vp, C.c_int32, C.c_int32, C.c_int32, _vp, _vp, _vp, C.c_size_t, _vp]),
    "vpk_model_forward_host": (C.c_int, [_vp, _vp, C.c_int32, C.c_int32, C.c_int32, _vp, _vp]),
    "vpk_model_forward_actions": (C.c_int, [_vp, _vp, _vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _vp, _vp, _vp,
                                            C.c_size_t, _vp]),
    "vpk_model_forward_host_actions": (C.c_int, [_vp, _vp, _vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _vp, _vp]),
    "vpk_model_microbatch": (C.c_int, [_vp, C.c_int32, C.POINTER(C.c_int32)]),
    "vpk_model_last_launch_count": (C.c_int, [_vp, C.POINTER(C.c_int64)]),
    "vpk_model_set_timing": (C.c_int, [_vp, C.c_int32]),
    "vpk_model_last_gemm_ms": (C.c_int, [_vp, C.POINTER(C.c_float), C.POINTER(C.c_int64), C.POINTER(C.c_double)]),
    "vpk_model_profile": (C.c_int, [_vp, C.c_char_p, C.c_size_t]),
    "vpk_model_destroy": (None, [_vp]),
    "vpk_metric_partial_sums": (C.c_int, [_vp, _vp, C.c_int32, C.c_int32, C.c_int64, _vp, _vp, _vp]),
    "vpk_metric_ssim_scratch_elems": (C.c_int64, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "vpk_metric_ssim_sums": (C.c_int, [_vp, _vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _vp, _vp, _vp]),
    "vpk_convlstm_cell_create": (C.c_int, [C.c_int32] * 8 + [_vp, _vp, C.POINTER(_vp)]),
    "vpk_convlstm_cell_step": (C.c_int, [_vp, C.c_int32] + [_vp] * 9),
    "vpk_convlstm_cell_backward": (C.c_int, [_vp, C.c_int32] + [_vp] * 11),
    "vpk_convlstm_cell_backward_peep": (C.c_int, [_vp, C.c_int32] + [_vp] * 17),
    "vpk_stlstm_cell_create": (C.c_int, [C.c_int32] * 7 + [_vp] * 5 + [C.POINTER(_vp)]),
    "vpk_stlstm_cell_set_layer_norm": (C.c_int, [_vp] * 9),
    "vpk_stlstm_cell_step": (C.c_int, [_vp, C.c_int32] + [_vp] * 10),
    "vpk_stlstm_ac_cell_create": (C.c_int, [C.c_int32] * 7 + [C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp)]),
    "vpk_stlstm_ac_cell_set_layer_norm": (C.c_int, [_vp, C.POINTER(_vp)]),
    "vpk_stlstm_ac_cell_step": (C.c_int, [_vp, C.c_int32] + [_vp] * 11),
    "vpk_causal_lstm_cell_create": (C.c_int, [C.c_int32] * 8 + [C.POINTER(_vp), C.POINTER(_vp)]),
    "vpk_causal_lstm_cell_step": (C.c_int, [_vp, C.c_int32] + [_vp] * 8),
    "vpk_ghu_cell_create": (C.c_int, [C.c_int32] * 6 + [_vp, _vp, C.POINTER(_vp)]),
    "vpk_ghu_cell_step": (C.c_int, [_vp, C.c_int32] + [_vp] * 4),
    "vpk_phycell_cell_set_action_convs": (C.c_int, [_vp, C.c_int32] + [_vp] * 4),
    "vpk_phycell_cell_step_action": (C.c_int, [_vp, C.c_int32] + [_vp] * 5),
    "vpk_phycell_cell_create": (C.c_int, [C.c_int32] * 7 + [_vp] * 8 + [C.POINTER(_vp)]),
    "vpk_phycell_cell_step": (C.c_int, [_vp, C.c_int32] + [_vp] * 4),
    "vpk_cell_destroy": (None, [_vp]),
    "vpk_last_error": (C.c_char_p, []),
    "vpk_version": (C.c_char_p, []),
    "vpk_device_ok": (C.c_int, []),
}

_lib = None


class NativeError(RuntimeError):
    pass


def lib():
    """Loads libvpk.so once; raises if it has not been built (no Python/CPU fallback exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NativeError(f"{LIB_PATH} not found: build it with `make -C vp_suite_b200/csrc` "
                              f"(or __graft_entry__.build()); vp_suite_b200 has no fallback path")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(status):
    if status != 0:
        msg = lib().vpk_last_error().decode("utf-8", "replace")
        if status == 1:
            raise ValueError(msg)          # the reference raises ValueError / AttributeError for these
        raise NativeError(f"libvpk error {status}: {msg}")


def ptr(t):
    """Raw device/host pointer of a torch tensor (or None)."""
    return None if t is None else C.c_void_p(t.data_ptr())
