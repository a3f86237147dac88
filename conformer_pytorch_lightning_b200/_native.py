"""ctypes binding of libcfm_b200.so (the C-ABI declared in include/cfm_b200.h).

There is NO fallback: if the shared library is missing or a call fails this module
raises.  The library is built in-tree by ``__graft_entry__.build()`` /
``make -C conformer_pytorch_lightning_b200/csrc``.
"""
from __future__ import annotations

import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcfm_b200.so")

F32, BF16 = 0, 1
EPI_BIAS, EPI_BIAS_SILU, EPI_BIAS_GLU, EPI_RESIDUAL, EPI_BIAS_RELU = 0, 1, 2, 3, 4
ENGINE_AUTO, ENGINE_SIMT, ENGINE_TC = 0, 1, 2

# every symbol include/cfm_b200.h declares (tests check the .so exports all of them)
EXPORTS = [
    "cfm_abi_version", "cfm_last_error", "cfm_init", "cfm_launch_count", "cfm_kernel_launches", "cfm_layernorm", "cfm_gemm", "cfm_gemm_ln", "cfm_ffn",
    "cfm_attention", "cfm_relpos_keys", "cfm_dwconv", "cfm_bn_stats", "cfm_bn_apply_silu", "cfm_subsample_ws_bytes",
    "cfm_subsample_conv", "cfm_conv_module", "cfm_mhsa_out", "cfm_ffn_chain", "cfm_ctc_ws_bytes", "cfm_ctc_argmax", "cfm_gemm_ex", "cfm_l2_prefetch", "cfm_l2_prefetch_multi",
    "cfm_ln_fwd_train", "cfm_ln_bwd", "cfm_silu_dropout_fwd", "cfm_silu_dropout_bwd", "cfm_resid_dropout_add",
    "cfm_scale_dropout_bwd", "cfm_glu_fwd", "cfm_glu_bwd", "cfm_bn_silu_bwd", "cfm_dwconv_wgrad", "cfm_softmax_fwd",
    "cfm_softmax_bwd", "cfm_colsum", "cfm_adam_step", "cfm_ctc_loss_ws_bytes", "cfm_ctc_loss_fwd", "cfm_ctc_loss_bwd",
    "cfm_fbank_frames", "cfm_fbank_power", "cfm_fbank_log_cmvn", "cfm_cmvn",
    "cfm_joint_add_tanh", "cfm_joint_tanh_bwd", "cfm_rnnt_loss_ws_bytes", "cfm_rnnt_loss_fwd", "cfm_rnnt_loss_bwd",
]

_lib = None
_lock = threading.Lock()
_inited_devices = set()

_p, _i, _i64, _f = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float


def _declare(lib):
    lib.cfm_abi_version.restype = _i
    lib.cfm_last_error.restype = ctypes.c_char_p
    lib.cfm_init.argtypes = [_i]
    lib.cfm_launch_count.restype = _i64
    lib.cfm_kernel_launches.argtypes = [ctypes.c_char_p]
    lib.cfm_kernel_launches.restype = _i64
    lib.cfm_layernorm.argtypes = [_p, _i, _i, _p, _p, _p, _p, _p, _p, _i, _p, _f, _p]
    lib.cfm_gemm.argtypes = [_p, _i, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p, _f, _p, _i, _p]
    lib.cfm_gemm_ln.argtypes = [_p, _i, _p, _p, _p, _i, _i, _i, _i, _i, _f, _p, _p, _p, _p, _p, _p, _i, _p, _f, _i, _p]
    lib.cfm_ffn.argtypes = [_p, _i, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _f, _p, _p, _p, _p, _p, _i, _p, _f, _p, _i, _p]
    lib.cfm_ffn_chain.argtypes = ([_p, _i, _i, _i, _i] + [_p, _p, _p, _p, _f, _p, _p, _p, _p] * 2 +
                                  [_p, _p, _p, _p, _p, _p, _i, _f, _p, _i, _p])
    lib.cfm_ctc_ws_bytes.argtypes = [_i, _i, _i]
    lib.cfm_ctc_ws_bytes.restype = _i64
    lib.cfm_ctc_argmax.argtypes = [_p, _i, _p, _p, _i, _i, _i, _i, _p, _p, _p, _i, _p]
    lib.cfm_mhsa_out.argtypes = [_p, _i64, _i64, _p, _i64, _i64, _p, _i64, _i64, _i, _i, _i, _i, _p, _i64, _i64, _p, _f,
                                 _p, _p, _p, _i, _p, _p, _p, _p, _f, _p, _i, _p]
    lib.cfm_conv_module.argtypes = [_p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _p, _p, _p, _p, _f, _p, _p, _i, _p]
    lib.cfm_attention.argtypes = [_p, _i64, _i64, _p, _i64, _i64, _p, _i64, _i64, _p, _i, _i, _i, _i,
                                  _p, _i64, _i64, _p, _f, _i, _i, _p]
    lib.cfm_relpos_keys.argtypes = [_p, _i64, _i64, _p, _i64, _p, _p, _p, _p, _i, _i, _i, _i, _p]
    lib.cfm_dwconv.argtypes = [_p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p]
    lib.cfm_bn_stats.argtypes = [_p, _i, _i, _p, _p, _p]
    lib.cfm_bn_apply_silu.argtypes = [_p, _i, _i, _p, _p, _p, _p, _p, _i, _p]
    lib.cfm_subsample_ws_bytes.argtypes = [_i, _i, _i, _i]
    lib.cfm_subsample_ws_bytes.restype = _i64
    lib.cfm_subsample_conv.argtypes = [_p, _i, _i, _i, _p, _p, _p, _p, _i, _p, _p, _p]
    lib.cfm_gemm_ex.argtypes = [_p, _i, _i64, _i64, _i64, _p, _i, _i64, _i64, _i64, _p, _i, _i64, _i64, _i64, _i,
                                _i, _i, _i, _i, _i, _i, _f, _i, _i, _p]
    _u64 = ctypes.c_uint64
    lib.cfm_joint_add_tanh.argtypes = [_p, _p, _p, _i, _i, _i, _i, _i, _p]
    lib.cfm_joint_tanh_bwd.argtypes = [_p, _p, _p, _p, _i, _i, _i, _i, _i, _p]
    lib.cfm_rnnt_loss_ws_bytes.argtypes = [_i, _i, _i]
    lib.cfm_rnnt_loss_ws_bytes.restype = _i64
    lib.cfm_rnnt_loss_fwd.argtypes = [_p, _i64, _i, _i, _i, _i, _i, _p, _i, _p, _p, _p, _p, _i, _p]
    lib.cfm_rnnt_loss_bwd.argtypes = [_p, _i64, _i, _i, _i, _i, _i, _i, _p, _i, _p, _p, _p, _p, _f, _p, _i, _p]
    lib.cfm_fbank_frames.argtypes = [_p, _i64, _p, _p, _p, _i, _i, _f, _p]
    lib.cfm_fbank_power.argtypes = [_p, _i, _p, _i, _i64, _i, _p]
    lib.cfm_fbank_log_cmvn.argtypes = [_p, _p, _p, _p, _p, _i, _i, _i, _p]
    lib.cfm_cmvn.argtypes = [_p, _p, _p, _p, _i64, _i, _p]
    lib.cfm_l2_prefetch.argtypes = [_p, _i64, _i, _p]
    lib.cfm_l2_prefetch_multi.argtypes = [_p, _p, _i, _i, _p]
    lib.cfm_ln_fwd_train.argtypes = [_p, _i, _i, _p, _p, _p, _i, _p, _p, _p, _f, _p]
    lib.cfm_ln_bwd.argtypes = [_p, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _p]
    lib.cfm_silu_dropout_fwd.argtypes = [_p, _p, _i, _i, _i, _f, _p, _i, _p]
    lib.cfm_silu_dropout_bwd.argtypes = [_p, _p, _p, _p, _i, _i, _i, _f, _p, _i, _p]
    lib.cfm_resid_dropout_add.argtypes = [_p, _p, _p, _i, _i, _i, _f, _p, _f, _p, _i, _p]
    lib.cfm_scale_dropout_bwd.argtypes = [_p, _p, _p, _i, _i, _i, _f, _p, _f, _p, _i, _p]
    lib.cfm_glu_fwd.argtypes = [_p, _p, _i, _i, _i, _p]
    lib.cfm_glu_bwd.argtypes = [_p, _p, _p, _p, _i, _i, _i, _p]
    lib.cfm_bn_silu_bwd.argtypes = [_p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _p]
    lib.cfm_dwconv_wgrad.argtypes = [_p, _p, _p, _p, _i, _i, _i, _i, _i, _p]
    lib.cfm_softmax_fwd.argtypes = [_p, _p, _p, _p, _i64, _i64, _i, _i, _i, _i, _i, _i, _f, _p, _i, _p]
    lib.cfm_softmax_bwd.argtypes = [_p, _p, _p, _i, _i, _i, _i, _i, _i, _f, _p, _i, _p]
    lib.cfm_colsum.argtypes = [_p, _i64, _p, _i, _i, _i, _p]
    lib.cfm_adam_step.argtypes = [_p, _p, _p, _p, _p, _i64, ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_double, _i, _f, _p]
    lib.cfm_ctc_loss_ws_bytes.argtypes = [_i, _i, _i]
    lib.cfm_ctc_loss_ws_bytes.restype = _i64
    lib.cfm_ctc_loss_fwd.argtypes = [_p, _i64, _i, _i, _i, _p, _i, _p, _p, _p, _p, _i, _p]
    lib.cfm_ctc_loss_bwd.argtypes = [_p, _i64, _i, _i, _i, _i, _p, _i, _p, _p, _p, _p, _f, _p, _i, _p]
    for name in EXPORTS:
        fn = getattr(lib, name)
        if name not in ("cfm_last_error", "cfm_launch_count", "cfm_kernel_launches", "cfm_abi_version", "cfm_subsample_ws_bytes", "cfm_ctc_ws_bytes",
                        "cfm_ctc_loss_ws_bytes", "cfm_rnnt_loss_ws_bytes"):
            fn.restype = _i


def lib():
    """Load (once) and return the ctypes handle; raises if the extension is not built."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise RuntimeError(
                        f"{LIB_PATH} not found: the CUDA extension is not built. Run "
                        "`python -c 'import __graft_entry__ as g; g.build()'` (there is no CPU fallback).")
                handle = ctypes.CDLL(LIB_PATH)
                _declare(handle)
                if handle.cfm_abi_version() != 1:
                    raise RuntimeError("libcfm_b200.so ABI version mismatch")
                _lib = handle
    return _lib


def check(rc):
    if rc != 0:
        raise RuntimeError("cfm_b200: " + lib().cfm_last_error().decode("utf-8", "replace"))


def init(device_index):
    if device_index not in _inited_devices:
        check(lib().cfm_init(int(device_index)))
        _inited_devices.add(device_index)


def launch_count():
    return int(lib().cfm_launch_count())


def kernel_launches(name):
    """Launches of one kernel family ("ffn_fused", "mhsa_fused", ...) issued through the library so far."""
    return int(lib().cfm_kernel_launches(name.encode()))
