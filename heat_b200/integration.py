"""Binding for a real Heat installation: route Heat's own ``KMeans`` / ``cdist`` through libhkmeans.

    import heat as ht, heat_b200.integration as hki
    hki.install()            # idempotent; hki.uninstall() restores the reference code
    ht.cluster.KMeans(n_clusters=64, init=c0).fit(x)     # x: ht.DNDarray on a CUDA device, split=0 or None

Only the hot path is replaced, and only when it applies (CUDA device, float32/float64, 2-D, split in {0, None});
everything else falls through to the reference implementation untouched:

* ``heat.cluster.KMeans.fit``            (heat/cluster/kmeans.py:105-148)  -> fused ``hk_lloyd_step`` loop
* ``heat.cluster.KMeans._assign_to_cluster`` (heat/cluster/_kcluster.py:352-370) -> ``hk_assign``
  (``predict`` and ``fit_predict`` go through it)
* ``heat.spatial.distance._euclidian / _gaussian / _gaussian_fast / _manhattan / _manhattan_fast`` (distance.py:17-133) ->
  ``hk_pairwise`` on the local blocks (``rbf``, ``manhattan``, ``cdist`` without the expansion; rings untouched).
* ``heat.spatial.distance._euclidian_fast`` (heat/spatial/distance.py:32-44) -> ``hk_cdist`` on the local blocks, which
  accelerates ``cdist(quadratic_expansion=True)`` for every caller without touching ``_dist``'s split logic
  (the operator plug point, distance.py:209-227).

* ``heat.cluster.KMedians / KMedoids`` (``fit``, ``_assign_to_cluster``), ``BatchParallelKMeans / BatchParallelKMedians``
  (``fit``, ``predict``) and ``heat.classification.KNeighborsClassifier.predict`` (default metric) -> the standalone classes of
  this package (``hk_assign_l1``, ``hk_select_*``, ``hk_nearest_rows_l1``, ``hk_kmex_update``, ``hk_topk_rows``,
  ``hk_knn_vote``) on the same device tensors; initialisation stays the reference's code.

The per-iteration allreduce uses NCCL through ``torch.distributed``: the process group must map rank r of ``x.comm`` to
the same rank (one process per GPU, ``cuda:{rank % device_count}`` as in heat/core/devices.py:116-120).
"""
from __future__ import annotations

import numpy as np
import torch

from . import engine as _engine
from .communication import ProcessGroupCommunication

_ORIG = {}


def _eligible(x) -> bool:
    t = x.larray
    return t.is_cuda and t.dtype in (torch.float32, torch.float64) and t.dim() == 2 and x.split in (None, 0)


def _pg(x):
    comm = ProcessGroupCommunication()
    if x.comm.size != comm.size or x.comm.rank != comm.rank:
        raise RuntimeError("heat_b200.integration: torch.distributed must be initialised with the same rank/size as x.comm")
    return comm


def install() -> bool:
    """Patch Heat in place.  Returns False (and does nothing) when Heat is not importable."""
    try:
        import heat as ht
        from heat.core.dndarray import DNDarray
        from heat.spatial import distance as hdist
    except Exception:
        return False
    if _ORIG:
        return True
    KM = ht.cluster.KMeans
    _ORIG.update(fit=KM.fit, assign=KM._assign_to_cluster, efast=hdist._euclidian_fast)

    def _wrap(t, gshape, dtype, split, like):
        return DNDarray(t, gshape, dtype, split, like.device, like.comm, True)

    def fit(self, x, oversampling=2, iter_multiplier=1):
        if not isinstance(x, DNDarray) or not _eligible(x):
            return _ORIG["fit"](self, x, oversampling, iter_multiplier)
        init = self.init
        if isinstance(init, DNDarray) and init.larray.dtype != x.larray.dtype:
            return _ORIG["fit"](self, x, oversampling, iter_multiplier)  # mixed dtypes: reference path (decided before init runs)
        if int(self.max_iter) < 1:
            return _ORIG["fit"](self, x, oversampling, iter_multiplier)
        self._initialize_cluster_centers(x, oversampling, iter_multiplier)  # reference code (init runs once)
        c0 = self._cluster_centers.larray
        if c0.dtype != x.larray.dtype:
            c0 = c0.to(x.larray.dtype)  # string initialisers sample rows of x: same dtype in practice
        xl = x.larray if x.larray.stride(1) == 1 else x.larray.contiguous()
        dev = xl.device
        eng = _engine.get_engine(dev)
        distributed = x.split is not None and x.comm.size > 1
        if distributed:
            eng.init_comm(_pg(x))
        row_ws = eng.row_workspace(xl.shape[0])
        c = c0.to(dev).contiguous().clone()
        c_prev = torch.empty_like(c)
        shift2 = torch.zeros((), dtype=c.dtype, device=dev)
        state = torch.zeros(4, dtype=torch.int32, device=dev)
        use_tol = self.tol is not None
        tol_cmp = float(np.float32(self.tol)) if use_tol else 0.0
        done, chunk = 0, (8 if use_tol else self.max_iter)
        while done < self.max_iter:
            todo = min(chunk, self.max_iter - done)
            eng.lloyd_run(xl, c, c_prev, use_tol, tol_cmp, shift2, state, distributed, todo, row_ws=row_ws)
            done += todo
            if use_tol and done < self.max_iter and int(state[0].item()):
                break
        self._n_iter = int(state[1].item())
        labels = torch.empty((xl.shape[0], 1), dtype=torch.int64, device=dev)
        eng.assign(xl, c_prev, labels, row_ws=row_ws)
        self._cluster_centers = _wrap(c, tuple(c.shape), self._cluster_centers.dtype, None, x)
        self._inertia = _wrap(shift2, (), self._cluster_centers.dtype, None, x)
        self._labels = DNDarray(labels, (x.shape[0], 1), ht.int64, x.split, x.device, x.comm, x.balanced)
        return self

    def _assign_to_cluster(self, x, eval_functional_value=False):
        c = self._cluster_centers.larray
        if not _eligible(x) or c.dtype != x.larray.dtype:
            return _ORIG["assign"](self, x, eval_functional_value)
        xl = x.larray if x.larray.stride(1) == 1 else x.larray.contiguous()
        dev = xl.device
        eng = _engine.get_engine(dev)
        labels = torch.empty((xl.shape[0], 1), dtype=torch.int64, device=dev)
        fv = torch.zeros(1, dtype=torch.float64, device=dev) if eval_functional_value else None
        eng.assign(xl, c.to(dev).contiguous(), labels, fv)
        if eval_functional_value:
            if x.split is not None and x.comm.size > 1:
                eng.init_comm(_pg(x))
                eng.allreduce_f64(fv)
            self._functional_value = _wrap(fv[0].to(xl.dtype), (), x.dtype, None, x)
        return DNDarray(labels, (x.shape[0], 1), ht.int64, x.split, x.device, x.comm, x.balanced)

    def _euclidian_fast(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        if not (x.is_cuda and y.is_cuda and x.dtype == y.dtype and x.dtype in (torch.float32, torch.float64)
                and x.dim() == 2 and y.dim() == 2):
            return _ORIG["efast"](x, y)
        xc = x if x.stride(1) == 1 else x.contiguous()
        yc = y if y.stride(1) == 1 else y.contiguous()
        out = torch.empty((xc.shape[0], yc.shape[0]), dtype=x.dtype, device=x.device)
        _engine.get_engine(x.device).cdist(xc, yc, out, quadratic_expansion=True, sqrt=True)
        return out

    def _tile(name, metric, expand):
        """A replacement for one of distance.py's tile metrics (:17-133): the same local blocks through hk_pairwise."""
        orig = getattr(hdist, name)
        _ORIG["tile_" + name] = orig

        def tile(x: torch.Tensor, y: torch.Tensor, sigma: float = 1.0) -> torch.Tensor:
            if not (x.is_cuda and y.is_cuda and x.dtype == y.dtype and x.dtype in (torch.float32, torch.float64)
                    and x.dim() == 2 and y.dim() == 2):
                return orig(x, y, sigma) if metric == "gaussian" else orig(x, y)
            xc = x if x.stride(1) == 1 else x.contiguous()
            yc = y if y.stride(1) == 1 else y.contiguous()
            out = torch.empty((xc.shape[0], yc.shape[0]), dtype=x.dtype, device=x.device)
            _engine.get_engine(x.device).pairwise(xc, yc, out, metric, expand, sigma)
            return out

        tile.__name__ = name
        setattr(hdist, name, tile)

    # ---- other consumers: Heat's own KMedians / KMedoids / BatchParallel* / KNeighborsClassifier objects,
    # the work delegated to the standalone classes of this package on views of the same device tensors
    from . import classification as _hbk
    from . import cluster as _hbc
    from .dndarray import DNDarray as _HB

    def _to_hb(a):
        return _HB(a.larray, tuple(a.shape), a.larray.dtype, a.split, a.larray.device, _pg(a), a.balanced)

    def _l1_fit(cls_name, mine_cls):
        cls = getattr(ht.cluster, cls_name)
        _ORIG["patch_" + cls_name + ".fit"] = (cls, "fit", cls.__dict__.get("fit"))
        _ORIG["patch_" + cls_name + "._assign_to_cluster"] = (cls, "_assign_to_cluster", cls.__dict__.get("_assign_to_cluster"))
        orig_fit = cls.fit
        orig_assign = cls._assign_to_cluster

        def _mine(self):
            kw = dict(n_clusters=self.n_clusters, init=_to_hb(self._cluster_centers), max_iter=self.max_iter,
                      random_state=self.random_state)
            if mine_cls is _hbc.KMedians:
                kw["tol"] = self.tol
            m = mine_cls(**kw)
            m._cluster_centers = _to_hb(self._cluster_centers)
            return m

        def l1_fit(self, x, oversampling=2, iter_multiplier=1):
            if not isinstance(x, DNDarray) or not _eligible(x) or int(self.max_iter) < 1:
                return orig_fit(self, x, oversampling, iter_multiplier)
            self._initialize_cluster_centers(x, oversampling, iter_multiplier)  # reference code
            hdt = self._cluster_centers.dtype
            m = _mine(self)
            m.fit(_to_hb(x))
            self._cluster_centers = _wrap(m.cluster_centers_.larray, tuple(m.cluster_centers_.shape), hdt, None, x)
            self._labels = DNDarray(m.labels_.larray, (x.shape[0], 1), ht.int64, x.split, x.device, x.comm, x.balanced)
            self._n_iter = m.n_iter_
            if m.inertia_ is not None:
                self._inertia = _wrap(m.inertia_.larray, (), hdt, None, x)
            return self

        def l1_assign(self, x, eval_functional_value=False):
            if not isinstance(x, DNDarray) or not _eligible(x):
                return orig_assign(self, x, eval_functional_value)
            m = _mine(self)
            lab = m._assign_to_cluster(_to_hb(x), eval_functional_value)
            if eval_functional_value:
                self._functional_value = _wrap(m.functional_value_.larray, (), x.dtype, None, x)
            return DNDarray(lab.larray, (x.shape[0], 1), ht.int64, x.split, x.device, x.comm, x.balanced)

        cls.fit = l1_fit
        cls._assign_to_cluster = l1_assign

    _l1_fit("KMedians", _hbc.KMedians)
    _l1_fit("KMedoids", _hbc.KMedoids)

    def _bp(cls_name):
        cls = getattr(ht.cluster, cls_name)
        base = cls.__mro__[1]  # _BatchParallelKCluster holds fit / predict
        if "patch_bp.fit" not in _ORIG:
            _ORIG["patch_bp.fit"] = (base, "fit", base.__dict__.get("fit"))
            _ORIG["patch_bp.predict"] = (base, "predict", base.__dict__.get("predict"))
            orig_fit, orig_predict = base.fit, base.predict

            def _mine(self):
                return _hbc._BatchParallelKCluster(self._p, self.n_clusters, self._init, self.max_iter, self.tol,
                                                   self.random_state, self.n_procs_to_merge)

            def bp_fit(self, x):
                if not isinstance(x, DNDarray) or not _eligible(x) or x.split != 0 or self._p not in (1, 2):
                    return orig_fit(self, x)
                m = _mine(self).fit(_to_hb(x))
                c = m.cluster_centers_.larray
                self._cluster_centers = DNDarray(c, tuple(c.shape), x.dtype, None, x.device, x.comm, True)
                self._n_iter = m.n_iter_
                return self

            def bp_predict(self, x):
                if (not isinstance(x, DNDarray) or not _eligible(x) or x.split != 0 or self._cluster_centers is None
                        or self._p not in (1, 2) or x.shape[1] != self._cluster_centers.shape[1]):
                    return orig_predict(self, x)
                m = _mine(self)
                m._cluster_centers = _to_hb(self._cluster_centers)
                lab = m.predict(_to_hb(x))
                self._functional_value = m.functional_value_
                return DNDarray(lab.larray, (x.shape[0], 1), ht.int32, x.split, x.device, x.comm, x.balanced)

            base.fit = bp_fit
            base.predict = bp_predict

    _bp("BatchParallelKMeans")

    KNN = ht.classification.kneighborsclassifier.KNeighborsClassifier
    _ORIG["patch_knn.predict"] = (KNN, "predict", KNN.__dict__.get("predict"))
    orig_knn_predict = KNN.predict

    def knn_predict(self, x):
        ok = (isinstance(x, DNDarray) and _eligible(x) and isinstance(self.x, DNDarray) and _eligible(self.x)
              and self.effective_metric_ is ht.spatial.cdist and self.x.larray.dtype == x.larray.dtype
              and self.y.larray.dim() == 2 and self.y.split in (None, 0))
        if not ok:
            return orig_knn_predict(self, x)
        m = _hbk.KNeighborsClassifier(n_neighbors=self.n_neighbors)
        m.x, m.y = _to_hb(self.x), _to_hb(self.y)
        cls = m.predict(_to_hb(x))
        self.classes_ = DNDarray(cls.larray, (x.shape[0],), ht.int64, cls.split, x.device, x.comm, x.balanced)
        return self.classes_

    KNN.predict = knn_predict

    KM.fit = fit
    KM._assign_to_cluster = _assign_to_cluster
    hdist._euclidian_fast = _euclidian_fast
    # rbf / manhattan / cdist without the expansion: _dist (layouts, rings) stays the reference's
    for name, metric, expand in (("_euclidian", "euclidean", False), ("_gaussian", "gaussian", False),
                                 ("_gaussian_fast", "gaussian", True), ("_manhattan", "manhattan", False),
                                 ("_manhattan_fast", "manhattan", True)):
        _tile(name, metric, expand)
    return True


def uninstall() -> None:
    if not _ORIG:
        return
    import heat as ht
    from heat.spatial import distance as hdist

    ht.cluster.KMeans.fit = _ORIG["fit"]
    ht.cluster.KMeans._assign_to_cluster = _ORIG["assign"]
    hdist._euclidian_fast = _ORIG["efast"]
    for key, fn in _ORIG.items():
        if key.startswith("tile_"):
            setattr(hdist, key[5:], fn)
        elif key.startswith("patch_"):
            owner, name, orig = fn
            if orig is None:
                delattr(owner, name)  # the attribute was inherited before the patch
            else:
                setattr(owner, name, orig)
    _ORIG.clear()
