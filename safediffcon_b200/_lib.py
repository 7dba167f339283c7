"""ctypes binding of libsafediffcon_b200.so (the C ABI declared in include/safediffcon_b200*.h).

There is no CPU fallback: if the library is missing, or a compute entry point is called with non-CUDA
tensors, this module raises instead of computing something else.
"""
import ctypes
import os

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SDC_LIB_PATH") or os.path.join(_PKG, "libsafediffcon_b200.so")   # override: A/B of two builds (scripts/)
_lib = None

c_f = ctypes.c_float
c_d = ctypes.c_double
c_i = ctypes.c_int
c_i64 = ctypes.c_int64
c_u64 = ctypes.c_uint64
c_p = ctypes.c_void_p


class StepCoef(ctypes.Structure):
    _fields_ = [("c1", c_f), ("c2", c_f), ("k_x0", c_f), ("k_eps", c_f), ("k_noise", c_f), ("sched", c_f),
                ("is_last", ctypes.c_int32), ("t", ctypes.c_int32)]


class Guidance(ctypes.Structure):
    _fields_ = [("mode", ctypes.c_int32), ("Q", c_f), ("u_bound_sq", c_f), ("w_score", c_f), ("scaler", c_f),
                ("nt", ctypes.c_int32)]


class ChainState(ctypes.Structure):
    _fields_ = [("step", ctypes.c_int32), ("reserved", ctypes.c_int32), ("seed", c_u64), ("sample_offset", c_i64)]


_SIGS = {
    "sdc_version": (c_i, []),
    "sdc_last_error": (ctypes.c_char_p, []),
    "sdc_launch_count": (c_i64, []),
    "sdc_burgers_nonfinite_rollouts": (c_i, [c_i, ctypes.POINTER(c_i64)]),
    "sdc_burgers_solve_free": (c_i, [c_p, c_p, c_p, c_i64, c_i, c_i, c_d, c_d, c_d, c_i, c_p]),
    "sdc_burgers_solve_cartesian": (c_i, [c_p, c_p, c_p, c_i64, c_i64, c_i, c_i, c_d, c_d, c_d, c_i, c_p]),
    "sdc_burgers_score": (c_i, [c_p, c_p, c_f, c_i64, c_i, c_i, c_p, c_p, c_p, c_p, c_p]),
    "sdc_burgers_control_score": (c_i, [c_p, c_i, c_p, c_f, c_p, c_i64, c_i, c_i, c_d, c_d, c_d, c_i, c_p, c_p, c_p, c_p, c_p]),
    "sdc_reverse_step": (c_i, [c_i, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_p, ctypes.POINTER(Guidance), c_p, c_p, c_p, c_p,
                               c_i, c_i, c_i, c_u64, c_i64, c_i64, c_i, c_i, c_p]),
    "sdc_write_conditions": (c_i, [c_p, c_p, c_p, c_p, c_i, c_i, c_i64, c_i, c_i, c_p]),
    "sdc_fill_normal": (c_i, [c_p, c_i64, c_i64, c_u64, c_i64, ctypes.c_int32, c_p]),
    "sdc_advance_counter": (c_i, [c_p, c_p]),
    "sdc_chain_state_set": (c_i, [c_p, ctypes.c_int32, c_u64, c_i64, c_p, c_i, c_p, c_i64, c_p]),
    "sdc_chain_state_advance": (c_i, [c_p, c_p, c_i, c_p, c_i64, c_p]),
    "sdc_reverse_step_state": (c_i, [c_i, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, ctypes.POINTER(Guidance), c_p, c_p, c_p, c_p,
                                     c_i, c_i, c_i, c_i64, c_i, c_i, c_p]),
    "sdc_count_launches": (None, [c_i64]),
    "sdc_burgers_fields": (c_i, [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_i64, c_i64, c_i, c_i, c_i, c_d, c_i, c_f, c_p]),
    "sdc_dataset_states": (c_i, [c_p, c_p, c_p, c_i64, c_i, c_i, c_i, c_i, c_f, c_i, c_p]),
    "sdc_safety_stat": (c_i, [c_p, c_p, c_i, c_f, c_i, c_i64, c_i, c_i, c_p]),
    "sdc_conformal_scores": (c_i, [c_p, c_p, c_p, c_p, ctypes.POINTER(Guidance), c_f, c_i64, c_i, c_i, c_p]),
    "sdc_normalize_weights": (c_i, [c_p, c_p, c_p, c_i64, c_p]),
    "sdc_kth_select_workspace": (c_i64, [c_i64]),
    "sdc_kth_select": (c_i, [c_p, c_i64, c_i64, c_p, c_p, c_p, c_p]),
}


def register(sigs):
    """Other modules (unet) add their entry points here before the first lib() call or after it."""
    _SIGS.update(sigs)
    if _lib is not None:
        _apply(_lib, sigs)


def _apply(l, sigs):
    for name, (res, args) in sigs.items():
        fn = getattr(l, name)
        fn.restype = res
        fn.argtypes = args


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"safediffcon_b200: CUDA library not built ({LIB_PATH} missing). Run `python -m safediffcon_b200.build` "
                "(needs nvcc; sm_100a). There is no CPU fallback.")
        l = ctypes.CDLL(LIB_PATH)
        _apply(l, _SIGS)
        _lib = l
    return _lib


def check(status):
    if status != 0:
        msg = lib().sdc_last_error().decode("utf-8", "replace")
        if status == 1:
            raise ValueError(f"safediffcon_b200: {msg}")
        raise RuntimeError(f"safediffcon_b200 (status {status}): {msg}")


def dev_f32(t, name="tensor"):
    """Contiguous fp32 CUDA view of t (copying only if needed); loud failure for CPU tensors."""
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise RuntimeError(f"safediffcon_b200: {name} is on {t.device}; this path is CUDA (sm_100a) only, no CPU fallback")
    if t.dtype != torch.float32:
        t = t.to(torch.float32)
    return t.contiguous()


def ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def stream_ptr(device=None):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def launch_count():
    return int(lib().sdc_launch_count())
