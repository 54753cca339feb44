"""ctypes binding of libgadpamsa.so (include/gadpamsa.h).

There is NO CPU fallback: if the shared library is missing, or a compute entry point is
called on a machine without a CUDA device, this module raises. PyTorch is used by the
callers only for device memory, streams and torch.distributed.
"""
from __future__ import annotations

import ctypes
import os

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, "libgadpamsa.so")

GAD_OK, GAD_EINVAL, GAD_ECUDA, GAD_ECAPACITY, GAD_ENOMEM = 0, -1, -2, -3, -4
MODE_ID = {"sp": 0, "cs": 1, "mo": 2}

c_u8p = ctypes.POINTER(ctypes.c_uint8)
c_i32p = ctypes.POINTER(ctypes.c_int32)
c_i64p = ctypes.POINTER(ctypes.c_int64)
c_vp = ctypes.c_void_p
c_int = ctypes.c_int
c_i64 = ctypes.c_int64

# name -> (restype, argtypes); every symbol include/gadpamsa.h declares
SIGNATURES = {
    "gad_version": (c_int, []),
    "gad_last_error": (ctypes.c_char_p, []),
    "gad_device_count": (c_int, [ctypes.POINTER(c_int)]),
    "gad_set_device": (c_int, [c_int]),
    "gad_malloc": (c_int, [ctypes.POINTER(c_vp), ctypes.c_size_t]),
    "gad_free": (c_int, [c_vp]),
    "gad_malloc_host": (c_int, [ctypes.POINTER(c_vp), ctypes.c_size_t]),
    "gad_free_host": (c_int, [c_vp]),
    "gad_memcpy_h2d": (c_int, [c_vp, c_vp, ctypes.c_size_t, c_vp]),
    "gad_memcpy_d2h": (c_int, [c_vp, c_vp, ctypes.c_size_t, c_vp]),
    "gad_memset": (c_int, [c_vp, c_int, ctypes.c_size_t, c_vp]),
    "gad_stream_sync": (c_int, [c_vp]),
    "gad_enable_peer_access": (c_int, [c_int]),
    "gad_fitness": (c_int, [c_vp, c_vp, c_i64, c_int, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "gad_fitness_pass": (c_int, [c_vp, c_vp, c_i64, c_int, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_int,
                                 c_vp, c_vp, c_int, c_vp, ctypes.c_size_t, c_vp]),
    "gad_debug_stamps": (c_int, [c_vp]),
    "gad_probe_read": (c_int, [c_vp, c_i64, c_int, c_int, c_int, c_vp, c_vp]),
    "gad_fitness_host": (c_int, [c_vp, c_vp, c_i64, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_int]),
    "gad_window_score_host": (c_int, [c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                      c_i64p, c_i32p]),
    "gad_workspace_bytes": (c_int, [c_i64, ctypes.POINTER(ctypes.c_size_t)]),
    "gad_rank_keys": (c_int, [c_vp, c_vp, c_vp, c_i64, c_int, c_vp, c_vp, c_vp, ctypes.c_size_t, c_vp]),
    "gad_rank_order": (c_int, [c_vp, c_i64, c_vp, c_vp, ctypes.c_size_t, c_vp]),
    "gad_select": (c_int, [c_vp, c_vp, c_vp, c_i64, c_int, c_i64, c_vp, c_vp, c_vp, c_vp, ctypes.c_size_t, c_vp]),
    "gad_compact": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_int, c_int, c_vp, c_vp, c_vp, c_vp,
                            c_vp, c_vp]),
    "gad_hof_update": (c_int, [c_vp, c_vp, c_vp, c_vp, c_i64, c_int, c_int, c_int, c_vp, c_vp, c_int, c_vp,
                               ctypes.c_size_t, c_vp]),
    "gad_hof_argbest": (c_int, [c_vp, c_vp, c_vp, c_i64, c_int, c_vp, c_vp, c_vp, ctypes.c_size_t, c_vp]),
    "gad_crossover_from": (c_int, [c_vp, c_vp, c_vp, c_vp, c_i64, c_vp, c_i64, c_int, c_int, c_int, c_vp]),
    "gad_crossover_peer": (c_int, [c_vp, c_vp, c_i64, c_vp, c_vp, c_i64, c_vp, c_i64, c_int, c_int, c_int, c_vp]),
    "gad_crossover": (c_int, [c_vp, c_vp, c_i64, c_vp, c_i64, c_int, c_int, c_int, c_vp]),
    "gad_pass_scatter": (c_int, [c_vp, c_i64, c_i64, c_i64, c_i64, c_i64, c_vp, c_vp, c_vp, c_vp, c_i64, c_int, c_int, c_vp]),
    "gad_pass_decide": (c_int, [c_vp, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64, c_int, c_i64, c_int, c_int, c_vp, c_int,
                                c_int, c_vp, c_int, c_int, c_vp, ctypes.c_size_t, c_vp]),
    "gad_mt_crossover_plan_host": (c_int, [c_vp, c_i64, c_int, c_i64, c_vp, c_vp]),
    "gad_mt_mutation_draws_host": (c_int, [c_vp, c_i64, ctypes.c_double, c_vp, c_vp, c_vp]),
    "gad_mt_crossover_plan_prob_host": (c_int, [c_vp, c_i64, c_int, c_i64, c_vp, ctypes.c_double, c_int, c_vp]),
    "gad_mt_population_host": (c_int, [c_vp, c_vp, c_vp, c_int, c_vp, c_int, c_i64, c_i64, c_vp, c_vp, c_int]),
    "gad_mt_population_shard_host": (c_int, [c_vp, c_vp, c_vp, c_int, c_vp, c_int, c_i64, c_i64, c_i64, c_i64, c_vp, c_vp,
                                             c_int]),
    "gad_population_device": (c_int, [c_vp, c_vp, c_vp, c_int, c_vp, c_int, c_i64, c_i64, c_i64, c_i64, c_vp, c_vp, c_int, c_vp]),
    "gad_worst_tiles": (c_int, [c_vp, c_vp, c_int, c_int, c_vp, c_i64, c_int, c_int, c_int, c_int, c_int, c_int,
                                c_vp, c_vp]),
    "gad_qnet_create": (c_int, [ctypes.POINTER(c_vp), c_int, c_int, c_int, c_int, c_int, c_int,
                                ctypes.POINTER(c_vp)]),
    "gad_qnet_destroy": (c_int, [c_vp]),
    "gad_qnet_dims": (c_int, [c_vp, ctypes.POINTER(c_int), ctypes.POINTER(c_int)]),
    "gad_qnet_forward": (c_int, [c_vp, c_vp, c_i64, ctypes.c_float, c_vp, c_vp, c_vp]),
    "gad_qnet_forward_host": (c_int, [c_vp, c_vp, c_i64, ctypes.c_float, c_vp, c_vp]),
    "gad_qnet_time_l1": (c_int, [c_vp, c_i64, c_int, ctypes.POINTER(ctypes.c_float),
                                 ctypes.POINTER(ctypes.c_double), c_vp]),
    "gad_mutate": (c_int, [c_vp, c_vp, c_vp, c_int, c_int, c_vp, c_vp, c_i64, ctypes.c_float, c_vp, c_i64p, c_vp]),
    "gad_mutate_stats": (c_int, [c_vp, c_i64p]),
    "gad_mutate_windows": (c_int, [c_vp, c_vp, c_int, c_int, c_vp, c_vp, c_i64, c_vp, c_vp]),
    "gad_windows_rep": (c_int, [c_vp, c_i64, c_int, c_vp, c_vp, c_i64, c_vp, c_vp]),
    "gad_episodes_run": (c_int, [c_vp, c_vp, c_vp, c_i64, c_i64, ctypes.c_float, c_vp, c_vp, c_vp, c_i64p, c_vp]),
    "gad_splice_results": (c_int, [c_vp, c_vp, c_vp, c_int, c_int, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
}


class GadError(RuntimeError):
    pass


class GadCapacityError(GadError):
    pass


_lib = None


def lib():
    """Load the library (once). Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise GadError(
                "libgadpamsa.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `python ga-dpamsa_b200/build_native.py`; there is no CPU fallback." % LIB_PATH)
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)          # AttributeError if the symbol is missing
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc: int) -> None:
    """Map a C status to the Python exception the reference would raise."""
    if rc == GAD_OK:
        return
    msg = lib().gad_last_error().decode("utf-8", "replace")
    if rc == GAD_EINVAL:
        raise ValueError(msg)
    if rc == GAD_ECAPACITY:
        raise GadCapacityError(msg)
    if rc == GAD_ENOMEM:
        raise MemoryError(msg)
    raise GadError(msg)


def call(name: str, *args) -> None:
    check(getattr(lib(), name)(*args))


_device_checked = False


def require_device() -> None:
    """Fail loudly when there is no GPU: the product path never computes on the CPU."""
    global _device_checked
    if _device_checked:
        return
    n = c_int(0)
    rc = lib().gad_device_count(ctypes.byref(n))
    if rc != GAD_OK or n.value < 1:
        raise GadError("no CUDA device visible: GA-DPAMSA-B200 has no CPU fallback (%s)"
                       % lib().gad_last_error().decode("utf-8", "replace"))
    _device_checked = True


def ptr(t) -> int:
    """data pointer of a torch tensor / numpy array (or None)."""
    if t is None:
        return None
    if hasattr(t, "data_ptr"):
        return t.data_ptr()
    return t.ctypes.data


def current_stream_ptr() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream
