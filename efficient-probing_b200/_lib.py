"""ctypes binding of libep_b200.so (the C ABI in include/ep_b200.h).

PyTorch is used for device memory and streams only: every call passes raw device pointers and the
current stream handle.  There is no fallback: if the library is missing it is built with nvcc, and
if that fails or the device is not a B200 the call raises."""
import ctypes
import os
import threading

import torch

from . import build as _build

EP_DTYPE_BF16, EP_DTYPE_F32 = 0, 1
_lock = threading.Lock()
_lib = None

c_void_p, c_int, c_float, c_size_t, c_ll = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_size_t, ctypes.c_longlong

_SIGNATURES = {
    "ep_abi_version": (c_int, []),
    "ep_strerror": (ctypes.c_char_p, [c_int]),
    "ep_device_check": (c_int, []),
    "ep_set_kernel_mode": (c_int, [c_int]),
    "ep_last_kernel_family": (c_int, []),
    "ep_set_gemm_mode": (c_int, [c_int]),
    "ep_launch_count": (ctypes.c_ulonglong, []),
    "ep_set_debug": (c_int, [c_int]),
    "ep_timing_count": (c_int, []),
    "ep_timing_get": (c_int, [c_int, ctypes.c_char_p, c_int, ctypes.POINTER(c_float)]),
    "ep_timing_reset": (c_int, []),
    "ep_debug_trace": (c_int, [c_void_p, c_int]),
    "ep_set_sm_limit": (c_int, [c_int]),
    "ep_kernel_family_for": (c_int, [c_int] * 5),
    "ep_pooled_layout": (c_int, [c_int] * 6),
    "ep_workspace_bytes": (c_size_t, [c_int] * 5),
    "ep_fwd": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_float] + [c_int] * 5 +
               [c_void_p] * 6 + [c_void_p, c_size_t, c_void_p]),
    "ep_bwd": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_float] + [c_int] * 5 +
               [c_void_p] * 7 + [c_void_p] * 3 + [c_void_p, c_size_t, c_void_p]),
    "ep_fwd_ex": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_float] + [c_int] * 5 +
                  [c_void_p] * 6 + [c_void_p, c_size_t, c_void_p]),
    "ep_bwd_ex": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_float] + [c_int] * 5 +
                  [c_void_p] * 4 + [c_int] + [c_void_p] * 3 + [c_void_p] * 4 + [c_void_p, c_size_t, c_void_p]),
    "ep_bwd_proj": (c_int, [c_void_p] * 5 + [c_int] * 6 + [c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "ep_bwd_pool": (c_int, [c_void_p, c_int, c_void_p, c_float] + [c_int] * 5 + [c_void_p, c_void_p, c_void_p, c_void_p,
                                                                                   c_void_p, c_size_t, c_void_p]),
    "ep_attention_maps": (c_int, [c_void_p, c_int, c_void_p, c_float] + [c_int] * 4 + [c_void_p, c_void_p, c_size_t, c_void_p]),
    "ep_bn_fwd": (c_int, [c_void_p, c_int, c_int, c_float, c_float, c_int] + [c_void_p] * 7),
    "ep_bn_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "ep_linear_workspace_bytes": (c_size_t, [c_int] * 3),
    "ep_linear_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "ep_linear_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                              c_void_p, c_size_t, c_void_p]),
    "ep_ce_fwd_bwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_float, c_float, c_void_p, c_void_p, c_void_p, c_void_p]),
    "ep_adamw_step": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "ep_sgd_step": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "ep_lars_step": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    # ABI 2: operand copies written by their producers (the *_ops variants; trailing `ops` bitmask)
    "ep_refresh_operands": (c_int, [c_void_p, c_void_p, c_float] + [c_int] * 6 + [c_void_p, c_size_t, c_void_p, c_int,
                                                                                  c_void_p, c_size_t, c_int, c_void_p]),
    "ep_fwd_ops": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_float] + [c_int] * 5 +
                   [c_void_p] * 6 + [c_void_p, c_size_t, c_int, c_void_p]),
    "ep_bwd_proj_ops": (c_int, [c_void_p] * 5 + [c_int] * 6 + [c_void_p, c_void_p, c_void_p, c_size_t, c_int, c_void_p]),
    "ep_bn_fwd_ops": (c_int, [c_void_p, c_int, c_int, c_float, c_float, c_int] + [c_void_p] * 6 +
                      [c_int, c_void_p, c_size_t, c_int, c_void_p]),
    "ep_bn_bwd_ops": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p] + [c_int] * 5 +
                      [c_void_p, c_size_t, c_int, c_void_p]),
    "ep_linear_fwd_ops": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_size_t, c_int,
                                  c_void_p]),
    "ep_linear_bwd_ops": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                  c_void_p, c_size_t, c_int, c_void_p]),
    "ep_ce_fwd_bwd_ops": (c_int, [c_void_p, c_void_p, c_int, c_int, c_float, c_float, c_void_p, c_void_p, c_void_p,
                                  c_void_p, c_void_p, c_int, c_void_p, c_size_t, c_int, c_void_p]),
}
EP_OPS_WEIGHTS, EP_OPS_INPUT, EP_OPS_FP32, EP_OPS_NO_DW, EP_OPS_ONLY_DW = 1, 2, 4, 8, 16
EXPORTED_SYMBOLS = tuple(_SIGNATURES)


class EPError(RuntimeError):
    pass


def lib_path():
    return _build.LIB


def load(build_if_missing=True):
    """dlopen the library (building it first when it is absent or stale) and type its entry points."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = _build.build() if build_if_missing else _build.LIB
        lib = ctypes.CDLL(path)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError here == the ABI and the header disagree
            fn.restype, fn.argtypes = res, args
        if lib.ep_abi_version() != 2:
            raise EPError("libep_b200.so ABI version mismatch")
        _lib = lib
        return lib


def check(code, what):
    if code == 0:
        return
    msg = load().ep_strerror(code).decode()
    if code < 0:
        raise ValueError(f"{what}: {msg} (ep_status {code})")
    raise EPError(f"{what}: CUDA error {code}: {msg}")


def kernel_timings(reset=True):
    """[(kernel name, microseconds)] recorded since the last reset (needs ep_set_debug(32))."""
    lib = load()
    out = []
    buf = ctypes.create_string_buffer(32)
    us = c_float()
    for i in range(lib.ep_timing_count()):
        if lib.ep_timing_get(i, buf, 32, ctypes.byref(us)) == 0:
            out.append((buf.value.decode(), float(us.value)))
    if reset:
        lib.ep_timing_reset()
    return out


def ptr(t):
    return None if t is None else t.data_ptr()


def stream_ptr(device=None):
    return torch.cuda.current_stream(device).cuda_stream


def require_cuda(t, name):
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: efficient_probing_b200 has no CPU path "
                           f"(the sm_100a library is the only implementation)")


def x_dtype_code(x):
    if x.dtype == torch.bfloat16:
        return EP_DTYPE_BF16
    if x.dtype == torch.float32:
        return EP_DTYPE_F32
    raise TypeError(f"token tensor must be bfloat16 or float32, got {x.dtype}")
