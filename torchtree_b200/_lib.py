"""ctypes binding of the ttb200 C ABI (include/ttb200.h).

The CUDA library is the product: if it is missing or cannot be loaded this
module raises -- there is no CPU or PyTorch fallback.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_int32, c_int64, c_uint8, c_void_p

_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libttb200.so")

TTB2_HOST = 0
TTB2_DEVICE = 1
TTB2_FLAG_PREALLOC_GRAD = 1
TTB2_FLAG_FORCE_GENERIC = 2
TTB2_FLAG_FUSED = 4
TTB2_FLAG_NO_MMA = 8
TTB2_FLAG_CHERRY = 16
TTB2_FLAG_NO_GRAPH = 32
TTB2_FLAG_NO_CHERRY = 64


class Ttb2Config(ctypes.Structure):
    _fields_ = [
        ("tip_count", c_int32),
        ("pattern_count", c_int32),
        ("state_count", c_int32),
        ("category_count", c_int32),
        ("max_draws", c_int32),
        ("code_count", c_int32),
        ("device", c_int32),
        ("flags", c_int32),
    ]


class EngineError(RuntimeError):
    """Raised when a ttb200 call returns a non-zero status."""


_lib = None

# every symbol include/ttb200.h declares
EXPORTED_SYMBOLS = (
    "ttb2_create",
    "ttb2_set_postorder",
    "ttb2_destroy",
    "ttb2_set_stream",
    "ttb2_synchronize",
    "ttb2_loglik_mats",
    "ttb2_grad_mats",
    "ttb2_loglik_eigen",
    "ttb2_grad_eigen",
    "ttb2_grad_eigen_packed",
    "ttb2_packed_count",
    "ttb2_loglik_q",
    "ttb2_loglik_expm",
    "ttb2_get_eigen",
    "ttb2_site_loglik",
    "ttb2_get_mats",
    "ttb2_enable_timing",
    "ttb2_phase_ms",
    "ttb2_compress_patterns",
    "ttb2_heights_create",
    "ttb2_heights_destroy",
    "ttb2_heights_forward",
    "ttb2_heights_backward",
    "ttb2_coalescent_constant",
    "ttb2_coalescent_piecewise",
    "ttb2_launch_count",
    "ttb2_device_bytes",
    "ttb2_eval_serial",
    "ttb2_get_config",
    "ttb2_last_error",
    "ttb2_version",
)


def lib_path() -> str:
    return _LIB_PATH


def load():
    """Load libttb200.so (once) and declare the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise EngineError(
            "torchtree_b200: %s is missing. Build it with `python -m torchtree_b200.build` "
            "(needs nvcc; sm_100a only). There is no CPU fallback." % _LIB_PATH
        )
    lib = ctypes.CDLL(_LIB_PATH, mode=ctypes.RTLD_LOCAL)
    vp = c_void_p
    lib.ttb2_version.restype = c_int32
    lib.ttb2_last_error.restype = c_char_p
    lib.ttb2_create.argtypes = [POINTER(Ttb2Config), vp, vp, vp, vp, POINTER(vp)]
    lib.ttb2_create.restype = c_int32
    lib.ttb2_set_postorder.argtypes = [vp, vp]
    lib.ttb2_set_postorder.restype = c_int32
    lib.ttb2_destroy.argtypes = [vp]
    lib.ttb2_destroy.restype = None
    lib.ttb2_set_stream.argtypes = [vp, vp]
    lib.ttb2_set_stream.restype = c_int32
    lib.ttb2_synchronize.argtypes = [vp]
    lib.ttb2_synchronize.restype = c_int32
    lib.ttb2_loglik_mats.argtypes = [vp, c_int32, vp, vp, c_int32, vp, c_int32, vp, c_int32]
    lib.ttb2_loglik_mats.restype = c_int32
    lib.ttb2_grad_mats.argtypes = [vp, vp, vp, vp, vp, c_int32]
    lib.ttb2_grad_mats.restype = c_int32
    lib.ttb2_loglik_eigen.argtypes = [
        vp, c_int32, vp, vp, c_int32, vp, c_int32, vp, vp, vp, c_int32, vp, c_int32, vp, c_int32]
    lib.ttb2_loglik_eigen.restype = c_int32
    lib.ttb2_loglik_q.argtypes = [
        vp, c_int32, vp, vp, c_int32, vp, c_int32, vp, c_int32, vp, c_int32, vp, c_int32]
    lib.ttb2_loglik_q.restype = c_int32
    lib.ttb2_loglik_expm.argtypes = [
        vp, c_int32, vp, vp, c_int32, vp, c_int32, vp, c_int32, vp, c_int32, vp, c_int32]
    lib.ttb2_loglik_expm.restype = c_int32
    lib.ttb2_get_eigen.argtypes = [vp, vp, vp, vp, c_int32]
    lib.ttb2_get_eigen.restype = c_int32
    lib.ttb2_grad_eigen.argtypes = [vp, vp, vp, vp, vp, vp, vp, c_int32]
    lib.ttb2_grad_eigen.restype = c_int32
    lib.ttb2_grad_eigen_packed.argtypes = [vp, vp, vp, c_int64, c_int32]
    lib.ttb2_grad_eigen_packed.restype = c_int32
    lib.ttb2_packed_count.argtypes = [vp]
    lib.ttb2_packed_count.restype = c_int64
    lib.ttb2_site_loglik.argtypes = [vp, vp, c_int32]
    lib.ttb2_site_loglik.restype = c_int32
    lib.ttb2_get_mats.argtypes = [vp, vp, c_int32]
    lib.ttb2_get_mats.restype = c_int32
    lib.ttb2_enable_timing.argtypes = [vp, c_int32]
    lib.ttb2_enable_timing.restype = c_int32
    lib.ttb2_phase_ms.argtypes = [vp, vp]
    lib.ttb2_phase_ms.restype = c_int32
    lib.ttb2_compress_patterns.argtypes = [vp, c_int32, c_int64, c_int32, vp, vp, vp]
    lib.ttb2_compress_patterns.restype = c_int32
    lib.ttb2_heights_create.argtypes = [c_int32, vp, vp, c_int32, ctypes.POINTER(vp)]
    lib.ttb2_heights_create.restype = c_int32
    lib.ttb2_heights_destroy.argtypes = [vp]
    lib.ttb2_heights_destroy.restype = c_int32
    lib.ttb2_heights_forward.argtypes = [vp, c_int32, vp, vp, c_int32]
    lib.ttb2_heights_forward.restype = c_int32
    lib.ttb2_heights_backward.argtypes = [vp, c_int32, vp, vp, vp, vp, c_int32]
    lib.ttb2_heights_backward.restype = c_int32
    lib.ttb2_coalescent_constant.argtypes = [c_int32, c_int32, c_int32, vp, vp, c_int32, vp, vp, vp,
                                             c_int32]
    lib.ttb2_coalescent_constant.restype = c_int32
    lib.ttb2_coalescent_piecewise.argtypes = [c_int32, c_int32, c_int32, vp, vp, c_int32, c_int32,
                                              vp, c_int32, vp, vp, vp, c_int32]
    lib.ttb2_coalescent_piecewise.restype = c_int32
    lib.ttb2_launch_count.argtypes = [vp]
    lib.ttb2_launch_count.restype = c_int64
    lib.ttb2_device_bytes.argtypes = [vp]
    lib.ttb2_device_bytes.restype = c_int64
    lib.ttb2_eval_serial.argtypes = [vp]
    lib.ttb2_eval_serial.restype = c_int64
    lib.ttb2_get_config.argtypes = [vp, POINTER(Ttb2Config)]
    lib.ttb2_get_config.restype = c_int32
    _lib = lib
    return lib


def check(status: int, what: str) -> None:
    if status != 0:
        msg = load().ttb2_last_error()
        raise EngineError(
            "%s failed (status %d): %s" % (what, status, msg.decode() if msg else "?")
        )
