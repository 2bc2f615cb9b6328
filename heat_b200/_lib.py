"""ctypes binding of ``libhkmeans.so`` (the C ABI declared in ``include/hkmeans.h``).

There is deliberately no fallback: if the shared library is missing, or a call fails, the caller
gets an exception.  PyTorch is used only for device memory, streams and ``torch.distributed``.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_int, c_int32, c_int64, c_void_p

HK_F32, HK_F64 = 0, 1
HK_LABEL_NONE, HK_LABEL_U8, HK_LABEL_I32, HK_LABEL_I64 = 0, 1, 2, 3
HK_METRIC_EUCLIDEAN, HK_METRIC_GAUSSIAN, HK_METRIC_MANHATTAN = 0, 1, 2
HK_PATH_AUTO, HK_PATH_SIMT, HK_PATH_TC, HK_PATH_GENERIC, HK_PATH_ROW128 = 0, 1, 2, 3, 4

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HK_LIB") or os.path.join(_HERE, "libhkmeans.so")  # HK_LIB: experiment builds only

#: every symbol include/hkmeans.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "hk_version": (c_int, []),
    "hk_last_error": (c_char_p, []),
    "hk_create": (c_int, [POINTER(c_void_p), c_int]),
    "hk_destroy": (c_int, [c_void_p]),
    "hk_chunk": (c_int, [c_int64, c_int, c_int, POINTER(c_int64), POINTER(c_int64)]),
    "hk_lloyd_accumulate": (
        c_int,
        [c_void_p, c_void_p, c_int64, c_int, c_int64, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p,
         c_void_p, c_int64, c_int, c_void_p],
    ),
    "hk_lloyd_finalize": (
        c_int,
        [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_double, c_void_p, c_void_p,
         c_void_p],
    ),
    "hk_lloyd_step": (
        c_int,
        [c_void_p, c_void_p, c_int64, c_int, c_int64, c_int, c_void_p, c_void_p, c_int, c_void_p, c_int,
         c_int, c_double, c_void_p, c_void_p, c_int, c_void_p, c_int64, c_int, c_void_p],
    ),
    "hk_lloyd_run": (
        c_int,
        [c_void_p, c_void_p, c_int64, c_int, c_int64, c_int, c_void_p, c_void_p, c_int, c_int, c_double,
         c_void_p, c_void_p, c_int, c_void_p, c_int64, c_int, c_int, c_void_p],
    ),
    "hk_row_ws_bytes": (c_int64, [c_int64]),
    "hk_assign": (
        c_int,
        [c_void_p, c_void_p, c_int64, c_int, c_int64, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p,
         c_void_p, c_int64, c_int, c_void_p],
    ),
    "hk_cdist": (
        c_int,
        [c_void_p, c_void_p, c_int64, c_int, c_int64, c_void_p, c_int64, c_int64, c_void_p, c_int64, c_int,
         c_int, c_int, c_void_p],
    ),
    "hk_pairwise": (
        c_int,
        [c_void_p, c_void_p, c_int64, c_int, c_int64, c_void_p, c_int64, c_int64, c_void_p, c_int64, c_int,
         c_int, c_int, c_double, c_void_p],
    ),
    "hk_assign_l1": (
        c_int,
        [c_void_p, c_void_p, c_int64, c_int, c_int64, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p],
    ),
    "hk_row_keep": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int64, c_int, c_void_p, c_void_p]),
    "hk_select_passes": (c_int, [c_int]),
    "hk_select_hist": (
        c_int,
        [c_void_p, c_void_p, c_int64, c_int, c_int64, c_int, c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p,
         c_void_p],
    ),
    "hk_select_step": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "hk_select_value": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "hk_kmex_update": (
        c_int,
        [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_double, c_double, c_void_p, c_void_p],
    ),
    "hk_nearest_rows_l1": (
        c_int,
        [c_void_p, c_void_p, c_int64, c_int, c_int64, c_int, c_void_p, c_int, c_int64, c_void_p, c_void_p, c_void_p],
    ),
    "hk_topk_rows": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "hk_knn_vote": (
        c_int,
        [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_int64, c_int, c_int64, c_int, c_void_p, c_void_p],
    ),
    "hk_comm_unique_id": (c_int, [c_void_p]),
    "hk_comm_init": (c_int, [c_void_p, c_int, c_int, c_void_p]),
    "hk_comm_destroy": (c_int, [c_void_p]),
    "hk_allreduce_f64": (c_int, [c_void_p, c_void_p, c_int64, c_void_p]),
    "hk_comm_peer_export": (c_int, [c_void_p, c_int64, c_void_p]),
    "hk_comm_peer_import": (c_int, [c_void_p, c_void_p]),
    "hk_comm_mode": (c_int, [c_void_p]),
    "hk_stats_read": (c_int, [c_void_p, POINTER(c_int64)]),
    "hk_graph_enable": (c_int, [c_void_p, c_int]),
    "hk_graph_launch_count": (c_int64, [c_void_p]),
    "hk_launch_count": (c_int64, [c_void_p]),
    "hk_last_variant": (c_char_p, [c_void_p]),
    "hk_profile_enable": (c_int, [c_void_p, c_int]),
    "hk_profile_read": (c_int, [c_void_p, POINTER(c_double), POINTER(c_int64)]),
}

_lib = None


class HKError(RuntimeError):
    """A libhkmeans call returned a non-zero status."""


def load() -> ctypes.CDLL:
    """Load libhkmeans.so (once) and attach prototypes.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing — build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(heat_b200 has no CPU or PyTorch fallback for the k-means hot path)"
        )
    lib = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the header and the library drifted apart
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().hk_last_error()
        raise HKError(f"{what} failed with status {rc}: {msg.decode() if msg else ''}")
