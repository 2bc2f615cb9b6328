"""Device engine: torch tensors in, C-ABI calls out.  One ``CudaEngine`` per (process, device)."""
from __future__ import annotations

import ctypes
import os
from typing import Optional

import torch

from . import _lib
from ._lib import (HK_F32, HK_F64, HK_LABEL_I32, HK_LABEL_I64, HK_LABEL_NONE, HK_LABEL_U8, HK_PATH_AUTO,
                   HK_METRIC_EUCLIDEAN, HK_METRIC_GAUSSIAN, HK_METRIC_MANHATTAN, HK_PATH_GENERIC, HK_PATH_ROW128,
                   HK_PATH_SIMT, HK_PATH_TC, check)

_DT = {torch.float32: HK_F32, torch.float64: HK_F64}
_LK = {torch.uint8: HK_LABEL_U8, torch.int32: HK_LABEL_I32, torch.int64: HK_LABEL_I64}
METRICS = {"euclidean": HK_METRIC_EUCLIDEAN, "gaussian": HK_METRIC_GAUSSIAN, "manhattan": HK_METRIC_MANHATTAN}
PATHS = {"auto": HK_PATH_AUTO, "simt": HK_PATH_SIMT, "tc": HK_PATH_TC, "generic": HK_PATH_GENERIC,
         "row128": HK_PATH_ROW128}


def _ptr(t: Optional[torch.Tensor]):
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def _stream(dev: torch.device):
    return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


class CudaEngine:
    """Owns an ``hk_handle_t``; every method enqueues work on torch's current stream and returns."""

    def __init__(self, device: torch.device):
        if device.type != "cuda":
            raise RuntimeError(
                "heat_b200 runs the k-means path on CUDA devices only (sm_100a); "
                f"got a tensor on {device}. There is no CPU fallback."
            )
        self.lib = _lib.load()
        self.device = device
        idx = device.index if device.index is not None else torch.cuda.current_device()
        self.index = idx
        h = ctypes.c_void_p()
        check(self.lib.hk_create(ctypes.byref(h), idx), "hk_create")
        self.h = h
        self.comm_size = 1
        self.comm_key = None

    def close(self):
        if getattr(self, "h", None):
            self.lib.hk_destroy(self.h)
            self.h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    # -- communicator ----------------------------------------------------------------------------------
    #: partial vectors of up to this many doubles go through the peer-memory mailbox (k*(d+1) <= 256K covers
    #: k=1024, d=128); longer ones use ncclAllReduce
    PEER_CAP = 1 << 18

    def init_comm(self, comm) -> None:
        """Create the communicator matching ``comm``: NCCL (id travels over torch.distributed) plus the
        peer-memory mailbox of the fused finish kernel (cudaIpc handles gathered over torch.distributed)."""
        key = (id(comm), comm.rank, comm.size)
        if key == self.comm_key:
            return
        ident = None
        if comm.rank == 0:
            buf = ctypes.create_string_buffer(128)
            check(self.lib.hk_comm_unique_id(buf), "hk_comm_unique_id")
            ident = buf.raw
        ident = comm.bcast_bytes(ident, root=0)
        check(self.lib.hk_comm_init(self.h, comm.size, comm.rank, ctypes.c_char_p(ident)), "hk_comm_init")
        self.comm_key = key
        self.comm_size = comm.size
        if comm.size > 1 and os.environ.get("HK_NO_PEER", "0") != "1":
            mine = ctypes.create_string_buffer(64)
            rc = self.lib.hk_comm_peer_export(self.h, self.PEER_CAP, mine)
            handles = comm.allgather_bytes(mine.raw if rc == 0 else b"\0" * 64)
            oks = comm.allgather_bytes(bytes([1 if rc == 0 else 0]))
            if all(o == b"\x01" for o in oks):
                rc = self.lib.hk_comm_peer_import(self.h, ctypes.c_char_p(b"".join(handles)))
            else:
                rc = 1
            oks = comm.allgather_bytes(bytes([1 if rc == 0 else 0]))  # also the barrier the mailbox needs
            if not all(o == b"\x01" for o in oks):
                # some rank could not map its peers: every rank drops to NCCL (the decision must be collective)
                check(self.lib.hk_comm_init(self.h, comm.size, comm.rank, ctypes.c_char_p(ident)), "hk_comm_init")

    def comm_mode(self) -> str:
        return {0: "single", 1: "nccl", 2: "peer"}[int(self.lib.hk_comm_mode(self.h))]

    def allreduce_f64(self, buf: torch.Tensor) -> None:
        assert buf.dtype == torch.float64 and buf.is_contiguous()
        check(self.lib.hk_allreduce_f64(self.h, _ptr(buf), buf.numel(), _stream(self.device)),
              "hk_allreduce_f64")

    def row_workspace(self, n_local: int) -> torch.Tensor:
        """Zeroed per-matrix workspace for ``row_ws=`` (tied to the CONTENT of one matrix: make a new one, or
        ``zero_()`` it, when the rows change)."""
        nbytes = int(self.lib.hk_row_ws_bytes(int(n_local)))
        return torch.zeros(max(nbytes, 4), dtype=torch.uint8, device=self.device)

    # -- hot path ----------------------------------------------------------------------------------------
    @staticmethod
    def _check_x(x: torch.Tensor, c: torch.Tensor):
        if x.dtype not in _DT:
            raise TypeError(f"unsupported dtype {x.dtype}")
        if c.dtype != x.dtype:
            raise TypeError("centroids must have the dtype of the data on the device path")
        if x.dim() != 2 or c.dim() != 2 or x.shape[1] != c.shape[1]:
            raise ValueError("shape mismatch between data and centroids")
        if x.stride(1) != 1 and x.shape[0] > 0:
            raise ValueError("rows of x must be contiguous")
        if not c.is_contiguous():
            raise ValueError("centroids must be contiguous")

    @staticmethod
    def _ws(row_ws):
        if row_ws is None:
            return ctypes.c_void_p(0), 0
        return ctypes.c_void_p(row_ws.data_ptr()), row_ws.numel() * row_ws.element_size()

    def lloyd_accumulate(self, x, c, partials, labels=None, path="auto", row_ws=None):
        self._check_x(x, c)
        n, d = x.shape
        ldx = x.stride(0) if n > 1 else d
        lk = HK_LABEL_NONE if labels is None else _LK[labels.dtype]
        wp, wb = self._ws(row_ws)
        check(self.lib.hk_lloyd_accumulate(self.h, _ptr(x), n, d, ldx, _DT[x.dtype], _ptr(c), c.shape[0],
                                           _ptr(labels), lk, _ptr(partials), wp, wb, PATHS[path],
                                           _stream(self.device)), "hk_lloyd_accumulate")

    def lloyd_finalize(self, partials, c_in, c_out, use_tol, tol_cmp, shift2, state):
        k, d = c_in.shape
        check(self.lib.hk_lloyd_finalize(self.h, _ptr(partials), _ptr(c_in), _ptr(c_out), k, d,
                                         _DT[c_in.dtype], int(use_tol), float(tol_cmp), _ptr(shift2),
                                         _ptr(state), _stream(self.device)), "hk_lloyd_finalize")

    def lloyd_step(self, x, c, c_prev, use_tol, tol_cmp, shift2, state, allreduce, labels=None, path="auto",
                   row_ws=None):
        self._check_x(x, c)
        n, d = x.shape
        ldx = x.stride(0) if n > 1 else d
        lk = HK_LABEL_NONE if labels is None else _LK[labels.dtype]
        wp, wb = self._ws(row_ws)
        check(self.lib.hk_lloyd_step(self.h, _ptr(x), n, d, ldx, _DT[x.dtype], _ptr(c), _ptr(c_prev),
                                     c.shape[0], _ptr(labels), lk, int(use_tol), float(tol_cmp),
                                     _ptr(shift2), _ptr(state), int(allreduce), wp, wb, PATHS[path],
                                     _stream(self.device)), "hk_lloyd_step")

    def lloyd_run(self, x, c, c_prev, use_tol, tol_cmp, shift2, state, allreduce, iters, path="auto", row_ws=None):
        """``iters`` Lloyd steps with one call (CUDA-graph replay inside the library)."""
        self._check_x(x, c)
        n, d = x.shape
        ldx = x.stride(0) if n > 1 else d
        wp, wb = self._ws(row_ws)
        check(self.lib.hk_lloyd_run(self.h, _ptr(x), n, d, ldx, _DT[x.dtype], _ptr(c), _ptr(c_prev), c.shape[0],
                                    int(use_tol), float(tol_cmp), _ptr(shift2), _ptr(state), int(allreduce), wp, wb,
                                    PATHS[path], int(iters), _stream(self.device)), "hk_lloyd_run")

    def assign(self, x, c, labels, fv=None, path="auto", row_ws=None):
        self._check_x(x, c)
        n, d = x.shape
        ldx = x.stride(0) if n > 1 else d
        wp, wb = self._ws(row_ws)
        check(self.lib.hk_assign(self.h, _ptr(x), n, d, ldx, _DT[x.dtype], _ptr(c), c.shape[0], _ptr(labels),
                                 _LK[labels.dtype] if labels is not None else HK_LABEL_NONE, _ptr(fv), wp, wb,
                                 PATHS[path], _stream(self.device)), "hk_assign")

    def cdist(self, x, y, out, quadratic_expansion: bool, sqrt: bool = True):
        m, f = x.shape
        n = y.shape[0]
        check(self.lib.hk_cdist(self.h, _ptr(x), m, f, x.stride(0) if m > 1 else f, _ptr(y), n,
                                y.stride(0) if n > 1 else f, _ptr(out), out.stride(0) if m > 1 else n,
                                _DT[x.dtype], int(quadratic_expansion), int(sqrt), _stream(self.device)),
              "hk_cdist")

    def pairwise(self, x, y, out, metric: str = "euclidean", expand: bool = False, sigma: float = 1.0):
        """out = metric(x, y) on local blocks; ``out`` may be a column slice of a wider matrix (row stride kept)."""
        m, f = x.shape
        n = y.shape[0]
        check(self.lib.hk_pairwise(self.h, _ptr(x), m, f, x.stride(0) if m > 1 else f, _ptr(y), n,
                                   y.stride(0) if n > 1 else f, _ptr(out), out.stride(0) if m > 1 else max(n, 1),
                                   _DT[x.dtype], METRICS[metric], int(expand), float(sigma), _stream(self.device)),
              "hk_pairwise")

    # -- other consumers of the assignment pattern (KMedians / KMedoids / kNN) ---------------------------------
    def assign_l1(self, x, c, labels, fv=None):
        self._check_x(x, c)
        n, d = x.shape
        check(self.lib.hk_assign_l1(self.h, _ptr(x), n, d, x.stride(0) if n > 1 else d, _DT[x.dtype], _ptr(c), c.shape[0],
                                    _ptr(labels), _LK[labels.dtype] if labels is not None else HK_LABEL_NONE, _ptr(fv),
                                    _stream(self.device)), "hk_assign_l1")

    def kmex_update(self, c, flag, atol: float, partials=None, medians=None, counts=None, rtol: float = 1e-5):
        k, d = c.shape
        check(self.lib.hk_kmex_update(self.h, _ptr(partials), _ptr(medians), _ptr(counts), _ptr(c), k, d, _DT[c.dtype],
                                      float(atol), float(rtol), _ptr(flag), _stream(self.device)), "hk_kmex_update")

    def cluster_medians(self, x, labels, k: int, allsum=None, drop_zero_rows: bool = True, lower: bool = False):
        """Medians of every (cluster, feature) over the rows of all ranks (hk_select_* protocol of include/hkmeans.h).
        ``allsum(t)`` sums an int64 device tensor over the ranks in place (None: one process).  Returns
        ``(medians [k, d], counts [k] int64)``; rows that are entirely zero are not counted (kmedians.py:76-79) unless
        ``drop_zero_rows`` is False."""
        n, d = x.shape
        dt, st = _DT[x.dtype], _stream(self.device)
        ldx = x.stride(0) if n > 1 else d
        dev = x.device
        if drop_zero_rows:
            keep = torch.empty(max(n, 1), dtype=torch.uint8, device=dev)
            check(self.lib.hk_row_keep(self.h, _ptr(x), n, d, ldx, dt, _ptr(keep), st), "hk_row_keep")
        else:
            keep = torch.ones(max(n, 1), dtype=torch.uint8, device=dev)
        lab = labels.reshape(-1)
        if lab.dtype != torch.int64 or not lab.is_contiguous():
            lab = lab.to(torch.int64).contiguous()
        prefix = torch.zeros((2, k, d), dtype=torch.int64, device=dev)  # uint64 bit patterns
        remaining = torch.zeros((2, k, d), dtype=torch.int64, device=dev)
        hist = torch.empty((2, k, d, 256), dtype=torch.int64, device=dev)
        counts = None
        for p in range(self.lib.hk_select_passes(dt)):
            hist.zero_()
            check(self.lib.hk_select_hist(self.h, _ptr(x), n, d, ldx, dt, _ptr(lab), _ptr(keep), k, _ptr(prefix), p,
                                          _ptr(hist), st), "hk_select_hist")
            if p == 0:
                hist[1].copy_(hist[0])  # pass 0 counts once for both targets
            if allsum is not None:
                allsum(hist)
            if p == 0:
                counts = hist[0, :, 0, :].sum(dim=1)  # kept rows per cluster
                remaining[0] = ((counts - 1).clamp(min=0) // 2).view(k, 1)
                remaining[1] = (counts // 2).view(k, 1)
            check(self.lib.hk_select_step(self.h, _ptr(hist), _ptr(remaining), _ptr(prefix), k, d, st), "hk_select_step")
        # lower = torch.median's convention (the lower of the two middle values); else ht.median's interpolation
        frac = torch.zeros(k, dtype=torch.float64, device=dev) if lower else torch.where(counts % 2 == 0, 0.5, 0.0).to(torch.float64)
        med = torch.empty((k, d), dtype=x.dtype, device=dev)
        check(self.lib.hk_select_value(self.h, _ptr(prefix), _ptr(frac), k, d, dt, _ptr(med), st), "hk_select_value")
        return med, counts

    def nearest_rows_l1(self, x, p, row_base: int):
        """(distance [k] float64, global row index [k] int64) of the shard row closest in L1 to every row of ``p``."""
        n, d = x.shape
        k = p.shape[0]
        bd = torch.full((k,), float("inf"), dtype=torch.float64, device=x.device)
        bi = torch.full((k,), torch.iinfo(torch.int64).max, dtype=torch.int64, device=x.device)
        if n > 0:
            check(self.lib.hk_nearest_rows_l1(self.h, _ptr(x), n, d, x.stride(0) if n > 1 else d, _DT[x.dtype], _ptr(p), k,
                                              int(row_base), _ptr(bd), _ptr(bi), _stream(self.device)),
                  "hk_nearest_rows_l1")
        return bd, bi

    def topk_rows(self, dmat, kk: int):
        m, n = dmat.shape
        vals = torch.empty((m, kk), dtype=dmat.dtype, device=dmat.device)
        idx = torch.empty((m, kk), dtype=torch.int64, device=dmat.device)
        check(self.lib.hk_topk_rows(self.h, _ptr(dmat), m, n, dmat.stride(0) if m > 1 else n, _DT[dmat.dtype], kk,
                                    _ptr(vals), _ptr(idx), _stream(self.device)), "hk_topk_rows")
        return vals, idx

    def knn_vote(self, idx, y):
        m, kk = idx.shape
        n, nc = y.shape
        out = torch.empty(m, dtype=torch.int64, device=idx.device)
        check(self.lib.hk_knn_vote(self.h, _ptr(idx), m, kk, _ptr(y), n, nc, y.stride(0) if n > 1 else nc, _DT[y.dtype],
                                   _ptr(out), _stream(self.device)), "hk_knn_vote")
        return out

    # -- introspection -------------------------------------------------------------------------------------
    def launch_count(self) -> int:
        return int(self.lib.hk_launch_count(self.h))

    def stats(self) -> dict:
        """Cold-path counters of the tensor-core passes since the previous call (synchronises)."""
        out = (ctypes.c_int64 * 6)()
        check(self.lib.hk_stats_read(self.h, out), "hk_stats_read")
        names = ("undecided_rows", "exact_pairs", "all_centroid_rows", "cold_warps", "rows", "passes")
        st = dict(zip(names, [int(v) for v in out]))
        st["undecided_frac"] = st["undecided_rows"] / st["rows"] if st["rows"] else 0.0
        return st

    def graph(self, enable: bool) -> None:
        check(self.lib.hk_graph_enable(self.h, int(enable)), "hk_graph_enable")

    def graph_launch_count(self) -> int:
        return int(self.lib.hk_graph_launch_count(self.h))

    def last_variant(self) -> str:
        return self.lib.hk_last_variant(self.h).decode()

    def profile(self, enable: bool) -> None:
        check(self.lib.hk_profile_enable(self.h, int(enable)), "hk_profile_enable")

    def profile_read(self):
        """(summed device ms of the dominant kernel, launches measured) since the last read."""
        ms, n = ctypes.c_double(), ctypes.c_int64()
        check(self.lib.hk_profile_read(self.h, ctypes.byref(ms), ctypes.byref(n)), "hk_profile_read")
        return ms.value, n.value


_ENGINES = {}


def get_engine(device: torch.device) -> CudaEngine:
    """Engine cache keyed by device (fails loudly on non-CUDA devices or a missing library)."""
    if device.type == "cuda" and device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    key = str(device)
    eng = _ENGINES.get(key)
    if eng is None:
        eng = CudaEngine(device)
        _ENGINES[key] = eng
    return eng


def set_engine_factory(factory) -> None:
    """Test hook: replace the engine constructor (used by the gloo host-logic tests, which inject a
    checker-backed engine defined under tests/).  Not used by the product path."""
    global get_engine
    _ENGINES.clear()

    def _get(device):
        key = str(device)
        if key not in _ENGINES:
            _ENGINES[key] = factory(device)
        return _ENGINES[key]

    import heat_b200.engine as me

    me.get_engine = _get
