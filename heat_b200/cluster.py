"""``KMeans`` with the reference's estimator API, running one fused CUDA pass + one NCCL allreduce per
Lloyd iteration.

Mirrors heat/cluster/kmeans.py:14-148 and heat/cluster/_kcluster.py:13-415 of the reference
(/root/reference): same constructor, ``fit`` / ``predict`` / ``fit_predict``, properties, exceptions and
behavioural quirks.
"""
from __future__ import annotations

from typing import Optional, Union

import numpy as np
import torch

from . import engine as _engine
from .communication import IN_PLACE, get_comm
from .dndarray import DNDarray, array

_FLOATS = (torch.float32, torch.float64)


class BaseEstimator:
    """get_params / set_params (reference: heat/core/base.py:37-93)."""

    _PARAMS = ("init", "max_iter", "n_clusters", "random_state", "tol")

    def get_params(self, deep: bool = True) -> dict:
        return {p: getattr(self, p) for p in self._PARAMS}

    def set_params(self, **params):
        if not params:
            return self
        valid = self.get_params()
        for key, value in params.items():
            if key not in valid:
                raise ValueError(f"Invalid parameter {key} for estimator {self}. "
                                 "Check the list of available parameters with `estimator.get_params().keys()`.")
            setattr(self, key, value)
        return self

    def __repr__(self, indent: int = 1) -> str:
        return f"{self.__class__.__name__}({self.get_params()})"


class KMeans(BaseEstimator):
    """K-Means clustering (Lloyd's algorithm) — drop-in for ``heat.cluster.KMeans``.

    Parameters are those of the reference (heat/cluster/kmeans.py:55-62).  ``init`` may be a DNDarray of
    shape (n_clusters, n_features) (the parity entry point, _kcluster.py:136-143), ``"random"`` or ``"kmeans++"`` /
    ``"probability_based"`` (k-means||, _kcluster.py:146-245, with the N-sized work on the device).
    ``"batchparallel"`` takes the centres of ``BatchParallelKMeans`` (_kcluster.py:249-275).
    """

    def __init__(self, n_clusters: int = 8, init: Union[str, DNDarray] = "random", max_iter: int = 300,
                 tol: float = 1e-4, random_state: Optional[int] = None):
        if isinstance(init, str) and init == "kmeans++":
            init = "probability_based"
        self.n_clusters = n_clusters
        self.init = init
        self.max_iter = max_iter
        self.tol = tol
        self.random_state = random_state
        self._cluster_centers = None
        self._functional_value = None
        self._labels = None
        self._inertia = None
        self._n_iter = None
        self._p = 2
        #: "auto" | "simt" | "tc" — kernel family (tests pin it; users leave it alone)
        self.kernel_path = "auto"
        #: iterations enqueued between two reads of the device-side convergence flag
        self.sync_every = 8

    # -- properties (reference: _kcluster.py:63-98) ------------------------------------------------------
    @property
    def cluster_centers_(self) -> DNDarray:
        return self._cluster_centers

    @property
    def labels_(self) -> DNDarray:
        return self._labels

    @property
    def inertia_(self):
        return self._inertia

    @property
    def n_iter_(self) -> int:
        return self._n_iter

    @property
    def functional_value_(self) -> DNDarray:
        return self._functional_value

    # -- initialisation (reference: _kcluster.py:100-281) -------------------------------------------------
    def _initialize_cluster_centers(self, x: DNDarray, oversampling: float, iter_multiplier: float):
        if not isinstance(x, DNDarray):
            raise ValueError(f"Input x needs to be a ht.DNDarray, but was {type(x)}")
        if oversampling < 2:
            raise ValueError(f"Oversampling factor should be at least 2, but was {oversampling}")
        if iter_multiplier < 1:
            raise ValueError(f"Iteration multiplier should be at least 1, but was {iter_multiplier}")
        # argument errors first (ValueError), layout limits second (NotImplementedError): the order in which the
        # reference reports them (tests/cluster/test_kmeans.py:68-100)
        if isinstance(self.init, DNDarray):
            if len(self.init.shape) != 2:
                raise ValueError(f"passed centroids need to be two-dimensional, but are {len(self.init.shape)}")
            if self.init.shape[0] != self.n_clusters or self.init.shape[1] != x.shape[1]:
                raise ValueError("passed centroids do not match cluster count or data shape")
        elif not (isinstance(self.init, str) and self.init in ("random", "probability_based", "batchparallel")):
            raise ValueError(
                'init needs to be one of "random", ht.DNDarray, "kmeans++", or "batchparallel", '
                f"but was {self.init}")
        if len(x.shape) != 2:
            raise NotImplementedError("Only 2D data matrices are currently supported")
        if x.split not in (None, 0):
            raise NotImplementedError("Not implemented for other splitting-axes")

        if isinstance(self.init, DNDarray):
            self._cluster_centers = self.init.resplit(None)
        elif self.init == "random":
            g = torch.Generator()
            g.manual_seed(0 if self.random_state is None else int(self.random_state))
            idx = torch.randint(0, max(x.shape[0] - 1, 1), (self.n_clusters,), generator=g)
            self._cluster_centers = _gather_rows(x, idx)
        elif self.init == "probability_based":
            self._cluster_centers = self._probability_based_init(x, oversampling, iter_multiplier)
        else:
            # init="batchparallel" (reference: _kcluster.py:249-275): the batch-parallel clusterer's centres
            if x.split != 0:
                raise NotImplementedError(
                    f"Batch parallel initalization only implemented for split = 0, but split was {x.split}")
            if self._p == 2:
                bp = BatchParallelKMeans(n_clusters=self.n_clusters, init="k-means++", max_iter=100,
                                         random_state=self.random_state)
            elif self._p == 1:
                bp = BatchParallelKMedians(n_clusters=self.n_clusters, init="k-medians++", max_iter=100,
                                           random_state=self.random_state)
            else:
                raise ValueError("Batch parallel initialization only implemented for KMeans and KMedians")
            bp.fit(x)
            self._cluster_centers = bp.cluster_centers_

    # -- k-means|| initialisation (reference: _kcluster.py:146-245, 283-350) ------------------------------------
    def _probability_based_init(self, x: DNDarray, oversampling: float, iter_multiplier: float) -> DNDarray:
        """``init="kmeans++"`` / ``"probability_based"``: the reference's k-means|| scheme with the device doing the
        N-sized work.  Same structure as the reference: one uniformly drawn row; ``int(iter_multiplier * log(cost))``
        rounds in which every row joins the candidate set with probability ``oversampling * d / sum(d)`` (d = distance
        to the nearest candidate so far; the reference samples on distances, not squared distances); candidates are
        weighted by the number of rows closest to them and reclustered to ``n_clusters`` centroids by weighted
        k-means++ seeding plus Lloyd iterations on the (small) candidate set.  Distances to the rows come from
        ``hk_cdist`` on the NEW candidates of a round only (a running minimum per row is kept), the weights from
        ``hk_assign``.  Random numbers come from torch generators seeded with ``random_state`` (Heat's own counter-based
        generator is not reproduced, so the sampled rows differ from the reference's for the same seed)."""
        import math
        import warnings

        xl, cdtype = _device_operands(x)
        dev = xl.device
        eng = _engine.get_engine(dev)
        comm = x.comm
        distributed = x.split is not None and comm.is_distributed()
        if distributed:
            eng.init_comm(comm)
        n_loc, d = xl.shape
        seed = 0 if self.random_state is None else int(self.random_state)
        g_all = torch.Generator().manual_seed(seed)  # identical stream on every rank: global decisions
        g_loc = torch.Generator(device=dev).manual_seed(seed * 7919 + 104729 * (comm.rank if distributed else 0) + 1)

        def allsum(v: torch.Tensor) -> torch.Tensor:
            if distributed:
                comm.Allreduce(IN_PLACE, v)
            return v

        def fold_min(dmin: torch.Tensor, new: torch.Tensor) -> None:
            # dmin[i] = min(dmin[i], min_j |x_i - new_j|): hk_cdist on row blocks that keep the temporary below 1 GiB
            m = new.shape[0]
            if m == 0 or n_loc == 0:
                return
            rows = max(1024, min(n_loc, (1 << 28) // max(m, 1)))
            buf = torch.empty((min(rows, n_loc), m), dtype=cdtype, device=dev)
            for r0 in range(0, n_loc, rows):
                r1 = min(n_loc, r0 + rows)
                out = buf[: r1 - r0]
                if self._p == 1:  # KMedians / KMedoids sample on Manhattan distances (their metric, kmedians.py:43-50)
                    eng.pairwise(xl[r0:r1], new, out, "manhattan", True)
                else:
                    eng.cdist(xl[r0:r1], new, out, quadratic_expansion=True)
                torch.minimum(dmin[r0:r1], out.min(dim=1).values, out=dmin[r0:r1])

        idx0 = torch.randint(0, max(x.shape[0] - 1, 1), (1,), generator=g_all)
        first = _gather_rows(x, idx0).larray.to(cdtype).contiguous()  # (1, d), replicated
        dmin0 = torch.full((n_loc,), float("inf"), dtype=cdtype, device=dev)
        fold_min(dmin0, first)
        cost0 = float(allsum(dmin0.sum(dtype=torch.float64).reshape(1))[0])
        num_iters = max(1, int(iter_multiplier * math.log(cost0))) if cost0 > 0 else 1

        def sample(ovs: float) -> torch.Tensor:
            cents, dmin = first, dmin0.clone()
            for _ in range(num_iters):
                total = float(allsum(dmin.sum(dtype=torch.float64).reshape(1))[0])
                prob = dmin * (ovs / max(total, 1e-12))
                pick = torch.rand(n_loc, generator=g_loc, device=dev, dtype=cdtype) <= prob
                local = xl[pick].reshape(-1, d)
                new = comm.Allgatherv_rows(local) if distributed else local
                if new.shape[0]:
                    new = new.contiguous()
                    cents = torch.cat([cents, new], dim=0)
                    fold_min(dmin, new)
            return cents

        cents = sample(float(oversampling))
        if cents.shape[0] <= self.n_clusters:
            warnings.warn(f"Oversampling={oversampling} is too low for data set."
                          "Increasing it by factor 10 automatically. And restarting centroid initialization.", UserWarning)
            oversampling = 10 * oversampling
            cents = sample(float(oversampling))
        if cents.shape[0] <= self.n_clusters:
            raise ValueError(f"The parameter oversampling={oversampling} and/or iter_multiplier={iter_multiplier} "
                             "are chosen too small for the initialization of cluster centers.")
        # weights = number of rows closest to each candidate
        m = cents.shape[0]
        lab = torch.empty(n_loc, dtype=torch.int32, device=dev)
        if self._p == 1:
            eng.assign_l1(xl, cents.contiguous(), lab)
        else:
            eng.assign(xl, cents.contiguous(), lab)
        weights = allsum(torch.bincount(lab.long(), minlength=m).to(torch.float64))
        # recluster the candidates (m x d, small): weighted k-means++ seeding + Lloyd, identically on every rank
        # (reference: batchparallelclustering.py:23-88 run on rank 0 and broadcast)
        xc, w = cents.double().cpu(), weights.cpu()
        k = self.n_clusters
        idxs = torch.zeros(k, dtype=torch.long)
        idxs[0] = torch.randint(0, m, (1,), generator=g_all)
        for i in range(1, k):
            dist = torch.cdist(xc, xc[idxs[:i]]).min(dim=1).values
            p = w * dist
            if float(p.sum()) <= 0:  # fewer distinct candidates than clusters
                p = torch.ones(m, dtype=torch.float64)
            idxs[i] = torch.multinomial(p, 1, generator=g_all)
        centers = xc[idxs].clone()
        tol = 0.0 if self.tol is None else float(self.tol)
        for _ in range(int(self.max_iter)):
            labels = torch.cdist(xc, centers).argmin(dim=1)
            old = centers.clone()
            for i in range(k):
                sel = labels == i
                if bool(sel.any()):  # an empty cluster keeps its centre
                    centers[i] = xc[sel].mean(dim=0)
            if torch.allclose(centers, old, atol=tol):
                break
        out = centers.to(device=dev, dtype=xl.dtype if x.larray.dtype in _FLOATS else cdtype)
        return DNDarray(out, (k, d), out.dtype, None, dev, comm, True)

    # -- the hot loop -------------------------------------------------------------------------------------
    def fit(self, x: DNDarray, oversampling: float = 2, iter_multiplier: float = 1):
        """Reference: heat/cluster/kmeans.py:105-148 — same results, one device pass per iteration."""
        if not isinstance(x, DNDarray):
            raise TypeError(f"Input needs to be a ht.DNDarray, but was {type(x)}")
        self._initialize_cluster_centers(x, oversampling, iter_multiplier)
        self._n_iter = 0

        xl, cdtype = _device_operands(x)
        dev = xl.device
        eng = _engine.get_engine(dev)
        distributed = x.split is not None and x.comm.is_distributed()
        if distributed:
            eng.init_comm(x.comm)

        centers0 = self._cluster_centers.larray.to(dev)
        c_dtype = centers0.dtype if centers0.dtype in _FLOATS else cdtype
        if c_dtype == torch.float64 and cdtype == torch.float32:
            # float32 data + float64 centroids: the reference promotes both operands of cdist to float64
            # (distance.py:392-395), so distances, sums and functional value are float64
            xl = xl.to(torch.float64)
            cdtype = torch.float64
        k, d = centers0.shape
        use_tol = self.tol is not None
        tol_cmp = float(np.float32(self.tol)) if use_tol else 0.0
        state = torch.zeros(4, dtype=torch.int32, device=dev)
        max_iter = int(self.max_iter)
        chunk = max_iter if not use_tol else max(1, int(self.sync_every))
        # per-matrix workspace of the tensor-core path (|x| bounds), tied to this fit's view of the rows
        row_ws = eng.row_workspace(xl.shape[0])

        if c_dtype == cdtype:
            # fused step: pass over the shard -> finish kernel (slot reduce, peer-memory sum over the ranks, finalize),
            # centroids updated in place; `chunk` iterations per call, replayed from a CUDA graph inside the library
            c = centers0.to(cdtype).contiguous().clone()
            c_prev = torch.empty_like(c)
            shift2 = torch.zeros((), dtype=cdtype, device=dev)
            done = 0
            while done < max_iter:
                todo = min(chunk, max_iter - done)
                eng.lloyd_run(xl, c, c_prev, use_tol, tol_cmp, shift2, state, distributed, todo,
                              path=self.kernel_path, row_ws=row_ws)
                done += todo
                if use_tol and done < max_iter:
                    if int(state[0].item()):  # one host sync per `sync_every` iterations
                        break
            st = state.cpu()
            self._n_iter = int(st[1])
            centers, pre = c, c_prev
        else:
            # centroids keep the dtype of `init` while distances run in the data's dtype (quirk Q6)
            c_lo = centers0.contiguous().clone()
            pre = torch.empty((k, d), dtype=cdtype, device=dev)
            part = torch.empty(k * (d + 1), dtype=torch.float64, device=dev)
            shift2 = torch.zeros((), dtype=c_dtype, device=dev)
            it = 0
            while it < max_iter:
                pre.copy_(c_lo)
                eng.lloyd_accumulate(xl, pre, part, path=self.kernel_path, row_ws=row_ws)
                if distributed:
                    eng.allreduce_f64(part)
                eng.lloyd_finalize(part, c_lo, c_lo, use_tol, tol_cmp, shift2, state)
                it += 1
                if use_tol and int(state[0].item()):
                    break
            self._n_iter = it
            centers = c_lo

        if self._n_iter == 0:  # max_iter == 0: the reference leaves labels undefined; keep it explicit
            raise ValueError("max_iter must be at least 1")

        # labels of the last iteration, i.e. against its pre-update centroids (quirk Q5)
        labels = torch.empty((xl.shape[0], 1), dtype=torch.int64, device=dev)
        eng.assign(xl, pre.to(cdtype), labels, path=self.kernel_path, row_ws=row_ws)

        comm = x.comm
        self._cluster_centers = DNDarray(centers, (k, d), centers.dtype, None, dev, comm, True)
        self._inertia = DNDarray(shift2, (), shift2.dtype, None, dev, comm, True)
        self._labels = DNDarray(labels, (x.shape[0], 1), torch.int64, x.split, dev, comm, x.balanced)
        return self

    def _assign_to_cluster(self, x: DNDarray, eval_functional_value: bool = False) -> DNDarray:
        """Reference: heat/cluster/_kcluster.py:352-370."""
        xl, cdtype = _device_operands(x)
        dev = xl.device
        eng = _engine.get_engine(dev)
        c = self._cluster_centers.larray.to(dev)
        if c.dtype == torch.float64 and cdtype == torch.float32:  # promotion of distance.py:392-395
            xl = xl.to(torch.float64)
            cdtype = torch.float64
        c = c.to(cdtype).contiguous()
        labels = torch.empty((xl.shape[0], 1), dtype=torch.int64, device=dev)
        fv = torch.zeros(1, dtype=torch.float64, device=dev) if eval_functional_value else None
        eng.assign(xl, c, labels, fv, path=self.kernel_path)
        if eval_functional_value:
            if x.split is not None and x.comm.is_distributed():
                eng.init_comm(x.comm)
                eng.allreduce_f64(fv)
            val = fv[0].to(cdtype)
            self._functional_value = DNDarray(val, (), cdtype, None, dev, x.comm, True)
        return DNDarray(labels, (x.shape[0], 1), torch.int64, x.split, dev, x.comm, x.balanced)

    def predict(self, x: DNDarray) -> DNDarray:
        """Reference: heat/cluster/_kcluster.py:398-415."""
        if not isinstance(x, DNDarray):
            raise ValueError(f"input needs to be a ht.DNDarray, but was  {type(x)}")
        return self._assign_to_cluster(x, eval_functional_value=True)

    def fit_predict(self, x: DNDarray) -> DNDarray:
        """Reference: heat/core/base.py:200-212."""
        self.fit(x)
        return self.predict(x)


def _device_operands(x: DNDarray):
    """Local shard as a row-contiguous float tensor + the arithmetic dtype (distance.py:392-403)."""
    xl = x.larray
    if xl.dtype not in _FLOATS:
        xl = xl.to(torch.float64 if xl.dtype == torch.int64 else torch.float32)
    if xl.dim() != 2:
        raise NotImplementedError("Only 2D data matrices are currently supported")
    if xl.shape[0] > 0 and xl.stride(1) != 1:
        xl = xl.contiguous()
    return xl, xl.dtype


def _gather_rows(x: DNDarray, idx: torch.Tensor) -> DNDarray:
    """Replicated copy of the global rows ``idx`` of a split=0 (or replicated) array."""
    k = idx.numel()
    d = x.shape[1]
    out = torch.zeros((k, d), dtype=x.larray.dtype, device=x.larray.device)
    if x.split is None:
        out.copy_(x.larray[idx.to(x.larray.device)])
    else:
        # true row offset of this shard: exclusive scan of the local row counts (arrays built with is_split=0
        # need not follow the balanced partition of comm.chunk)
        n_loc = x.larray.shape[0]
        counts = [int.from_bytes(b, "little") for b in x.comm.allgather_bytes(int(n_loc).to_bytes(8, "little"))]
        off = sum(counts[: x.comm.rank])
        for j, g in enumerate(idx.tolist()):
            if off <= g < off + n_loc:
                out[j] = x.larray[g - off]
        if x.comm.is_distributed():
            x.comm.Allreduce(IN_PLACE, out)
    return DNDarray(out, (k, d), out.dtype, None, out.device, x.comm, True)


class _L1Cluster(KMeans):
    """Shared part of KMedians / KMedoids: Manhattan metric (p = 1), assignment by ``hk_assign_l1``, medians of every
    cluster by distributed radix selection (``hk_select_*``), the reference's fit loops with their per-iteration
    convergence test on the host (heat/cluster/kmedians.py:105-147, kmedoids.py:112-156)."""

    _INIT_ALIAS = ""

    def __init__(self, n_clusters: int = 8, init: Union[str, DNDarray] = "random", max_iter: int = 300,
                 tol: Optional[float] = 1e-4, random_state: Optional[int] = None):
        if isinstance(init, str) and init == self._INIT_ALIAS:
            init = "probability_based"
        super().__init__(n_clusters=n_clusters, init=init, max_iter=max_iter, tol=tol, random_state=random_state)
        self._p = 1

    def _operands(self, x: DNDarray):
        xl, cdtype = _device_operands(x)
        eng = _engine.get_engine(xl.device)
        c = self._cluster_centers.larray.to(xl.device)
        if c.dtype == torch.float64 and cdtype == torch.float32:  # promotion of distance.py:392-395
            xl, cdtype = xl.to(torch.float64), torch.float64
        return xl, cdtype, eng, c.to(cdtype).contiguous()

    def _assign_to_cluster(self, x: DNDarray, eval_functional_value: bool = False) -> DNDarray:
        """Reference: heat/cluster/_kcluster.py:352-370 with metric = manhattan(expand=True), p = 1."""
        xl, cdtype, eng, c = self._operands(x)
        dev = xl.device
        labels = torch.empty((xl.shape[0], 1), dtype=torch.int64, device=dev)
        fv = torch.zeros(1, dtype=torch.float64, device=dev) if eval_functional_value else None
        eng.assign_l1(xl, c, labels, fv)
        if eval_functional_value:
            if x.split is not None and x.comm.is_distributed():
                x.comm.Allreduce(IN_PLACE, fv)
            self._functional_value = DNDarray(fv[0].to(cdtype), (), cdtype, None, dev, x.comm, True)
        return DNDarray(labels, (x.shape[0], 1), torch.int64, x.split, dev, x.comm, x.balanced)

    def _cluster_medians(self, x: DNDarray, labels: DNDarray):
        """(medians [k, d] on the device, kept-row counts [k] on the host) over all ranks."""
        xl, cdtype, eng, _ = self._operands(x)
        distributed = x.split is not None and x.comm.is_distributed()
        allsum = (lambda t: x.comm.Allreduce(IN_PLACE, t)) if distributed else None
        med, counts = eng.cluster_medians(xl, labels.larray, self.n_clusters, allsum)
        return med, counts.cpu()

    def _random_row(self, x: DNDarray) -> torch.Tensor:
        """Failsafe of the reference for a cluster without points: a uniformly drawn data row
        (kmedians.py:82-95; torch's generator instead of Heat's, so the row differs for the same seed)."""
        if not hasattr(self, "_failsafe_rng"):
            self._failsafe_rng = torch.Generator().manual_seed(0 if self.random_state is None else int(self.random_state))
        idx = torch.randint(0, x.shape[0], (1,), generator=self._failsafe_rng)
        return _gather_rows(x, idx).larray[0]

    def _check_fit_input(self, x):
        if not isinstance(x, DNDarray):
            raise ValueError(f"input needs to be a ht.DNDarray, but was {type(x)}")


class KMedians(_L1Cluster):
    """K-Medians clustering — drop-in for ``heat.cluster.KMedians`` (heat/cluster/kmedians.py:11-147): Manhattan
    metric, centroids = per-feature medians of the assigned rows."""

    _INIT_ALIAS = "kmedians++"

    def _update_centroids(self, x: DNDarray, matching_centroids: DNDarray) -> torch.Tensor:
        """Reference: heat/cluster/kmedians.py:60-103."""
        med, counts = self._cluster_medians(x, matching_centroids)
        new = self._cluster_centers.larray.clone()
        new.copy_(med.to(new.dtype))
        for j in (counts == 0).nonzero().view(-1).tolist():
            new[j] = self._random_row(x).to(new.dtype)
        return new

    def fit(self, x: DNDarray, oversampling: float = 2, iter_multiplier: float = 1):
        """Reference: heat/cluster/kmedians.py:105-147."""
        self._check_fit_input(x)
        self._initialize_cluster_centers(x, oversampling, iter_multiplier)
        self._n_iter = 0
        dev = _device_operands(x)[0].device
        c = self._cluster_centers.larray.to(dev)
        self._cluster_centers = DNDarray(c, tuple(c.shape), c.dtype, None, dev, x.comm, True)
        matching = None
        for _ in range(int(self.max_iter)):
            self._n_iter += 1
            matching = self._assign_to_cluster(x)
            new = self._update_centroids(x, matching)
            inertia = ((self._cluster_centers.larray - new) ** 2).sum()
            self._inertia = DNDarray(inertia, (), inertia.dtype, None, dev, x.comm, True)
            self._cluster_centers = DNDarray(new, tuple(new.shape), new.dtype, None, dev, x.comm, True)
            if self.tol is not None and float(inertia) <= float(np.float32(self.tol)):
                break
        if matching is None:
            raise ValueError("max_iter must be at least 1")
        self._labels = matching
        return self


class KMedoids(_L1Cluster):
    """K-Medoids clustering — drop-in for ``heat.cluster.KMedoids`` (heat/cluster/kmedoids.py:11-156): Manhattan
    metric, every centroid is the data row closest to the median of its cluster."""

    _INIT_ALIAS = "kmedoids++"

    def __init__(self, n_clusters: int = 8, init: Union[str, DNDarray] = "random", max_iter: int = 300,
                 random_state: Optional[int] = None):
        super().__init__(n_clusters=n_clusters, init=init, max_iter=max_iter, tol=0.0, random_state=random_state)

    def _update_centroids(self, x: DNDarray, matching_centroids: DNDarray) -> torch.Tensor:
        """Reference: heat/cluster/kmedoids.py:57-110."""
        med, counts = self._cluster_medians(x, matching_centroids)
        xl, cdtype, eng, _ = self._operands(x)
        comm = x.comm
        distributed = x.split is not None and comm.is_distributed()
        row_base = 0
        if distributed:
            row_base = sum(comm.row_counts(xl.shape[0])[: comm.rank])
        bd, bi = eng.nearest_rows_l1(xl, med.to(cdtype).contiguous(), row_base)
        if distributed:
            # (distance, global index) of every rank; the smallest distance wins, the lowest index on ties
            cand = comm.allgather_bytes(torch.stack([bd, bi.double()]).cpu().numpy().tobytes())
            both = torch.stack([torch.frombuffer(bytearray(b), dtype=torch.float64).view(2, -1) for b in cand])  # [p,2,k]
            dist, gidx = both[:, 0], both[:, 1]
            order = torch.argsort(dist + 0.0, dim=0, stable=True)  # ranks hold ascending index ranges
            best = order[0]
            idx = gidx.gather(0, best.view(1, -1)).view(-1).long()
        else:
            idx = bi.cpu()
        new = self._cluster_centers.larray.clone()
        ok = counts > 0
        if bool(ok.any()):
            rows = _gather_rows(x, idx[ok]).larray.to(new.dtype)
            new[ok.to(new.device)] = rows
        for j in (~ok).nonzero().view(-1).tolist():
            new[j] = self._random_row(x).to(new.dtype)
        return new

    def fit(self, x: DNDarray, oversampling: float = 2, iter_multiplier: float = 1):
        """Reference: heat/cluster/kmedoids.py:112-156."""
        self._check_fit_input(x)
        self._initialize_cluster_centers(x, oversampling, iter_multiplier)
        self._n_iter = 0
        dev = _device_operands(x)[0].device
        c = self._cluster_centers.larray.to(dev)
        self._cluster_centers = DNDarray(c, tuple(c.shape), c.dtype, None, dev, x.comm, True)
        matching = None
        for _ in range(int(self.max_iter)):
            self._n_iter += 1
            matching = self._assign_to_cluster(x)
            new = self._update_centroids(x, matching)
            if torch.equal(self._cluster_centers.larray, new):
                break
            self._cluster_centers = DNDarray(new, tuple(new.shape), new.dtype, None, dev, x.comm, True)
        if matching is None:
            raise ValueError("max_iter must be at least 1")
        self._labels = matching
        return self


# ---- batch-parallel clusterers (heat/cluster/batchparallelclustering.py) ----------------------------------------------
def _plus_plus_rows(xl: torch.Tensor, n_clusters: int, p: int, gen: Optional[torch.Generator], eng) -> torch.Tensor:
    """k-means++ / k-medians++ seeding of one shard (reference: _initialize_plus_plus, batchparallelclustering.py:23-49):
    first row uniform, every further row drawn with probability proportional to its distance to the nearest chosen row.
    The N-sized distance work runs on the device (``hk_pairwise`` against the newest row, running minimum); the draws use
    the same CPU generator calls as the reference, so a seeded run picks the same rows."""
    n = xl.shape[0]
    max_samples = 2**24 - 1  # torch.multinomial's category limit, as in the reference
    if n > max_samples:
        sub = torch.randint(0, n, (max_samples,), generator=gen)
        xl = xl[sub.to(xl.device)].contiguous()
        n = max_samples
    idxs = torch.zeros(n_clusters, dtype=torch.long)
    idxs[0] = torch.randint(0, n, (1,), generator=gen)
    dmin = torch.full((n,), float("inf"), dtype=xl.dtype, device=xl.device)
    col = torch.empty((n, 1), dtype=xl.dtype, device=xl.device)
    for i in range(1, n_clusters):
        eng.pairwise(xl, xl[idxs[i - 1]:idxs[i - 1] + 1].contiguous(), col, "manhattan" if p == 1 else "euclidean", False)
        torch.minimum(dmin, col.view(-1), out=dmin)
        idxs[i] = torch.multinomial(dmin.cpu(), 1, generator=gen)
    return xl[idxs.to(xl.device)].clone()


def _kmex(xl: torch.Tensor, p: int, n_clusters: int, init, max_iter: int, tol: float, random_state: Optional[int], eng):
    """Single-shard k-means (p = 2) / k-medians (p = 1) of the batch-parallel clusterers (reference: _kmex,
    batchparallelclustering.py:52-86): empty clusters keep their centre, convergence = allclose(new, old, atol=tol).
    One device pass per iteration (Lloyd pass, or L1 assignment + radix-selected lower medians) and one tiny update kernel;
    the flag is read once per iteration like the reference's ``torch.allclose``."""
    gen = torch.Generator().manual_seed(int(random_state)) if random_state is not None else None
    if isinstance(init, torch.Tensor):
        if tuple(init.shape) != (n_clusters, xl.shape[1]):
            raise ValueError("if a torch tensor, init must have shape (n_clusters, n_features).")
        centers = init.to(device=xl.device, dtype=xl.dtype).contiguous().clone()
    elif init == "++":
        centers = _plus_plus_rows(xl, n_clusters, p, gen, eng)
    elif init == "random":
        centers = xl[torch.randint(0, xl.shape[0], (n_clusters,), generator=gen).to(xl.device)].clone()
    else:
        raise ValueError("init must be a torch tensor with initial centers, string '++', or 'random'.")
    k, d = centers.shape
    flag = torch.zeros(1, dtype=torch.int32, device=xl.device)
    part = torch.empty(k * (d + 1), dtype=torch.float64, device=xl.device)
    labels = torch.empty(xl.shape[0], dtype=torch.int64, device=xl.device)
    it = 0
    for it in range(1, int(max_iter) + 1):
        if p == 1:
            eng.assign_l1(xl, centers, labels)
            med, counts = eng.cluster_medians(xl, labels, k, None, drop_zero_rows=False, lower=True)
            eng.kmex_update(centers, flag, tol, medians=med, counts=counts)
        else:
            eng.lloyd_accumulate(xl, centers, part)
            eng.kmex_update(centers, flag, tol, partials=part)
        if int(flag.item()):
            break
    return centers, it


class _BatchParallelKCluster(BaseEstimator):
    """Reference: _BatchParallelKCluster (batchparallelclustering.py:98-331): every rank clusters its own shard, the
    per-rank centres are merged by clustering them again (hierarchically, ``n_procs_to_merge`` at a time)."""

    _PARAMS = ("init", "max_iter", "n_clusters", "n_procs_to_merge", "random_state", "tol")

    def __init__(self, p: int, n_clusters: int, init: str, max_iter: int, tol: float, random_state, n_procs_to_merge):
        if not isinstance(n_clusters, int):
            raise TypeError(f"n_clusters must be int, but was {type(n_clusters)}")
        if n_clusters <= 0:
            raise ValueError(f"n_clusters must be positive, but was {n_clusters}")
        if not isinstance(max_iter, int):
            raise TypeError(f"max_iter must be int, but was {type(max_iter)}")
        if max_iter <= 0:
            raise ValueError(f"max_iter must be positive, but was {max_iter}")
        if not isinstance(tol, float):
            raise TypeError(f"tol must be float, but was {type(tol)}")
        if tol <= 0:
            raise ValueError(f"tol must be positive, but was {tol}")
        if not isinstance(random_state, int) and random_state is not None:
            raise TypeError(f"random_state must be int or None, but was {type(random_state)}")
        if not isinstance(n_procs_to_merge, int) and n_procs_to_merge is not None:
            raise TypeError(f"procs_to_merge must be int or None, but was {type(n_procs_to_merge)}")
        if n_procs_to_merge is not None and n_procs_to_merge <= 1:
            raise ValueError(f"If an integer, procs_to_merge must be > 1, but was {n_procs_to_merge}.")
        self.n_clusters = n_clusters
        self._init = init
        self.max_iter = max_iter
        self.tol = tol
        self.random_state = random_state
        self.n_procs_to_merge = n_procs_to_merge
        self._p = p
        self._cluster_centers = None
        self._n_iter = None
        self._functional_value = None

    @property
    def cluster_centers_(self) -> DNDarray:
        return self._cluster_centers

    @property
    def n_iter_(self) -> int:
        return self._n_iter

    @property
    def functional_value_(self):
        return self._functional_value

    @staticmethod
    def _check_input(x):
        if not isinstance(x, DNDarray):
            raise TypeError(f"input needs to be a ht.DNDarray, but was {type(x)}")
        if not x.ndim == 2:
            raise ValueError(f"input needs to be 2D, but was {x.ndim}D")
        if x.split != 0:
            raise ValueError(f"input needs to be split along the sample axis, but was split along {x.split}")

    def fit(self, x: DNDarray):
        """Reference: batchparallelclustering.py:171-262."""
        self._check_input(x)
        xl, _ = _device_operands(x)
        eng = _engine.get_engine(xl.device)
        comm = x.comm
        size, rank = (comm.size, comm.rank) if comm.is_distributed() else (1, 0)
        seed = None if self.random_state is None else self.random_state + rank
        centers, n_iters = _kmex(xl, self._p, self.n_clusters, self._init, self.max_iter, self.tol, seed, eng)
        merge = self.n_procs_to_merge if self.n_procs_to_merge is not None else size
        current = list(range(size))
        k, d = centers.shape
        while len(current) > 1:
            everyone = comm.Allgatherv_rows(centers).view(size, k, d)  # collective at every level, on all ranks
            if rank in current:
                pos = current.index(rank)
                if pos % merge == 0:  # root of its group: cluster the group's centres
                    members = current[pos:pos + merge]
                    if len(members) > 1:
                        gathered = everyone[members].reshape(-1, d).contiguous()
                        centers, extra = _kmex(gathered, self._p, self.n_clusters, self._init, self.max_iter, self.tol, seed, eng)
                        n_iters += extra
            current = [current[i] for i in range(len(current)) if i % merge == 0]
        if size > 1:
            centers = comm.Allgatherv_rows(centers).view(size, k, d)[0].clone()  # Bcast from rank 0
        self._cluster_centers = DNDarray(centers, (k, d), centers.dtype, None, xl.device, comm, True)
        self._n_iter = n_iters
        return self

    def predict(self, x: DNDarray) -> DNDarray:
        """Reference: batchparallelclustering.py:264-331 — int32 labels, functional value as a Python float."""
        self._check_input(x)
        if self._cluster_centers is None:
            raise RuntimeError("fit needs to be called before predict")
        if x.shape[1] != self._cluster_centers.shape[1]:
            raise ValueError(f"input needs to have {self._cluster_centers.shape[1]} features, but has {x.shape[1]}")
        xl, _ = _device_operands(x)
        eng = _engine.get_engine(xl.device)
        c = self._cluster_centers.larray.to(device=xl.device, dtype=xl.dtype).contiguous()
        labels = torch.empty((xl.shape[0], 1), dtype=torch.int32, device=xl.device)
        fv = torch.zeros(1, dtype=torch.float64, device=xl.device)
        if self._p == 1:
            eng.assign_l1(xl, c, labels, fv)
        else:
            eng.assign(xl, c, labels, fv)
        if x.comm.is_distributed():
            x.comm.Allreduce(IN_PLACE, fv)
        self._functional_value = float(fv.item())
        return DNDarray(labels, (x.shape[0], 1), torch.int32, x.split, xl.device, x.comm, x.balanced)


class BatchParallelKMeans(_BatchParallelKCluster):
    """Drop-in for ``heat.cluster.BatchParallelKMeans`` (batchparallelclustering.py:339-393)."""

    def __init__(self, n_clusters: int = 8, init: str = "k-means++", max_iter: int = 300, tol: float = 1e-4,
                 random_state: Optional[int] = None, n_procs_to_merge: Optional[int] = None):
        if not isinstance(init, str):
            raise TypeError(f"init must be str, but was {type(init)}")
        if init == "k-means++":
            _init = "++"
        elif init == "random":
            raise NotImplementedError("random initialization for batch parallel k-means is currently not supported due to "
                                      "instable behaviour of the algorithm. Use init='k-means++' instead.")
        else:
            raise ValueError(f"init must be 'k-means++' or 'random', but was {init}")
        super().__init__(2, n_clusters, _init, max_iter, tol, random_state, n_procs_to_merge)
        self.init = init


class BatchParallelKMedians(_BatchParallelKCluster):
    """Drop-in for ``heat.cluster.BatchParallelKMedians`` (batchparallelclustering.py:396-450)."""

    def __init__(self, n_clusters: int = 8, init: str = "k-medians++", max_iter: int = 300, tol: float = 1e-4,
                 random_state: Optional[int] = None, n_procs_to_merge: Optional[int] = None):
        if not isinstance(init, str):
            raise TypeError(f"init must be str, but was {type(init)}")
        if init == "k-medians++":
            _init = "++"
        elif init == "random":
            raise NotImplementedError("random initialization for batch parallel k-medians is currently not supported due to "
                                      "instable behaviour of the algorithm. Use init='k-medians++' instead.")
        else:
            raise ValueError(f"init must be 'k-medians++' or 'random', but was {init}")
        super().__init__(1, n_clusters, _init, max_iter, tol, random_state, n_procs_to_merge)
        self.init = init
