"""ctypes binding of the C ABI in include/hy2dl_b200.h (libhy2dl_b200.so, built in-tree for sm_100a).

There is no fallback: if the library is missing or a call fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "csrc", "libhy2dl_b200.so")

MAX_PAR = 16
N_STATES = 5
PREC_FP32, PREC_TF32X3, PREC_TF32 = 0, 1, 2
MODEL_SHM, MODEL_HBV, MODEL_LINRES, MODEL_NONSENSE = 0, 1, 2, 3
WS_LSTM_FWD, WS_LSTM_BWD, WS_LSTM_WGRAD = 0, 1, 2
LAYOUT_ROWMAJOR, LAYOUT_RI32 = 0, 1

p_f = C.c_void_p  # device pointers travel as integers
i64 = C.c_int64
i32 = C.c_int


class ParMap(C.Structure):
    """hy2dl_parmap (include/hy2dl_b200.h)."""
    _fields_ = [
        ("n_par", C.c_int32), ("n_members", C.c_int32), ("n_routing", C.c_int32), ("warmup", C.c_int32),
        ("dynamic", C.c_int32 * MAX_PAR), ("lo", C.c_float * (MAX_PAR + 2)), ("hi", C.c_float * (MAX_PAR + 2)),
        ("sim_off", C.c_int64 * (MAX_PAR + 2)),
    ]


class HeadFwd(C.Structure):
    """hy2dl_head_fwd (include/hy2dl_b200.h): optional fusion of the head into hy2dl_lstm_fwd."""
    _fields_ = [("w", C.c_void_p), ("bias", C.c_void_p), ("pm", C.POINTER(ParMap)), ("par_sim", C.c_void_p),
                ("par_warm", C.c_void_p)]


class HeadBwd(C.Structure):
    """hy2dl_head_bwd (include/hy2dl_b200.h): optional fusion of the head adjoint into hy2dl_lstm_bwd."""
    _fields_ = [("w", C.c_void_p), ("pm", C.POINTER(ParMap)), ("par_sim", C.c_void_p), ("dpar_sim", C.c_void_p),
                ("dz", C.c_void_p)]


class HeadWgrad(C.Structure):
    """hy2dl_head_wgrad (include/hy2dl_b200.h): optional head part of hy2dl_lstm_wgrad."""
    _fields_ = [("dz", C.c_void_p), ("t0", C.c_int64), ("n_columns", C.c_int32), ("dw", C.c_void_p), ("db", C.c_void_p)]


_SIGNATURES = {
    # name: (restype, argtypes)
    "hy2dl_abi_version": (i32, []),
    "hy2dl_last_error": (C.c_char_p, []),
    "hy2dl_check_device": (i32, [i32]),
    "hy2dl_workspace_bytes": (i64, [i32, i64, i64, i64, i64, i64]),
    "hy2dl_pack_x": (i32, [p_f, i64, i64, i64, i64, p_f, i32, p_f]),
    "hy2dl_pack_forcing": (i32, [p_f, i64, i64, i64, p_f, p_f]),
    "hy2dl_lstm_fwd": (i32, [p_f, i64, i64, i64, i64, i64, p_f, p_f, p_f, p_f, p_f, p_f, p_f, p_f, p_f, i64, i32,
                              C.POINTER(HeadFwd), p_f]),
    "hy2dl_lstm_bwd": (i32, [i64, i64, i64, p_f, p_f, p_f, p_f, i64, p_f, p_f, p_f, i64, i32, C.POINTER(HeadBwd), p_f]),
    "hy2dl_lstm_wgrad": (i32, [p_f, p_f, p_f, i64, i64, i64, i64, i64, p_f, p_f, p_f, p_f, i64, i32, C.POINTER(HeadWgrad), p_f]),
    "hy2dl_parmap_layout": (i64, [C.POINTER(ParMap), i64, i64]),
    "hy2dl_head_map_fwd": (i32, [p_f, i64, i64, i64, p_f, p_f, C.POINTER(ParMap), p_f, p_f, i32, p_f]),
    "hy2dl_head_map_bwd": (i32, [p_f, p_f, p_f, i64, i64, i64, p_f, C.POINTER(ParMap), p_f, p_f, p_f, p_f, i64, i32, p_f]),
    "hy2dl_conceptual_fwd": (i32, [i32, p_f, i64, i64, i64, i64, C.POINTER(p_f), C.POINTER(i64), p_f, p_f, p_f, p_f]),
    "hy2dl_conceptual_bwd": (i32, [i32, p_f, i64, i64, i64, i64, C.POINTER(p_f), C.POINTER(i64), p_f, p_f, p_f, p_f,
                                   C.POINTER(p_f), p_f, p_f]),
    "hy2dl_uh_fwd": (i32, [p_f, p_f, p_f, i64, i64, p_f, p_f, p_f]),
    "hy2dl_uh_bwd": (i32, [p_f, p_f, p_f, p_f, p_f, i64, i64, p_f, p_f, p_f, p_f]),
    "hy2dl_loss_acc_floats": (i64, []),
    "hy2dl_head_bwd_workspace_bytes": (i64, [i64]),
    "hy2dl_sqnorm_floats": (i64, []),
    "hy2dl_loss_nse_reduce": (i32, [p_f, p_f, p_f, i64, p_f, p_f]),
    "hy2dl_loss_nse_finish": (i32, [p_f, p_f, p_f, i64, p_f, p_f, p_f, p_f, p_f]),
    "hy2dl_loss_wrmse_reduce": (i32, [p_f, p_f, i64, p_f, p_f]),
    "hy2dl_loss_wrmse_finish": (i32, [p_f, p_f, i64, p_f, p_f, p_f, p_f, p_f]),
    "hy2dl_nse_basins": (i32, [p_f, p_f, i64, i64, p_f, p_f]),
    "hy2dl_window_gather": (i32, [p_f] * 8 + [i64] * 7 + [p_f] * 4 + [i32, p_f]),
    "hy2dl_grad_sqnorm": (i32, [p_f, i64, p_f, p_f]),
    "hy2dl_clip_adam": (i32, [p_f, p_f, p_f, p_f, i64, p_f, C.c_float, C.c_float, p_f, C.c_float, C.c_float, C.c_float,
                              p_f, p_f]),
}

_lock = threading.Lock()
_lib = None
_checked_devices = set()


def load() -> C.CDLL:
    """Load the library (once) and attach the prototypes. Raises if it has not been built."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"hy2dl_b200: {LIB_PATH} is missing. Build it with `python -m hy2dl_b200.build` "
                "(or __graft_entry__.build()); there is no CPU or PyTorch fallback for this path.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError here means header and library disagree
            fn.restype = res
            fn.argtypes = args
        if lib.hy2dl_abi_version() != 1:
            raise RuntimeError(f"hy2dl_b200: ABI version {lib.hy2dl_abi_version()} != 1")
        _lib = lib
        return lib


def exported_symbols():
    return sorted(_SIGNATURES)


def last_error() -> str:
    return load().hy2dl_last_error().decode()


def check(rc: int):
    if rc != 0:
        raise RuntimeError(f"hy2dl_b200 C-ABI call failed ({rc}): {last_error()}")


def require_cuda(t: torch.Tensor, name: str = "tensor") -> torch.Tensor:
    """No CPU fallback: every tensor that reaches a kernel must be fp32 on an sm_100 device."""
    if not t.is_cuda:
        raise RuntimeError(f"hy2dl_b200: {name} is on {t.device}; this path runs on CUDA (sm_100a) only, "
                           "there is no CPU fallback")
    if t.dtype != torch.float32:
        raise RuntimeError(f"hy2dl_b200: {name} has dtype {t.dtype}; only float32 is supported")
    dev = t.device.index if t.device.index is not None else torch.cuda.current_device()
    if dev not in _checked_devices:
        check(load().hy2dl_check_device(dev))
        _checked_devices.add(dev)
    return t


def ptr(t):
    """Device pointer of a contiguous tensor (None -> NULL)."""
    if t is None:
        return None
    assert t.is_contiguous(), "internal: non-contiguous tensor handed to a kernel"
    return t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


# launch counter: bench.py reports how many of OUR kernels launches were issued (host-side count of
# C-ABI calls times the launches each performs is kept simple: one increment per C-ABI call)
calls = 0


def call(name: str, *args):
    global calls
    calls += 1
    check(getattr(load(), name)(*args))
