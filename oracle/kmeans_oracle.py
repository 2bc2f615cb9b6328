"""TEST INFRASTRUCTURE — CPU restatement of Heat's k-means Lloyd path.

This module is the *checker*.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it; nothing
under ``heat_b200/`` does.  It restates, call for call, what the reference
(helmholtz-analytics/heat, v1.9.0-dev, mounted at /root/reference) computes on
this path, using the same third-party arithmetic the reference uses (torch CPU:
``torch.mm``, ``torch.min``, ``clamp``, ``sqrt`` — pinned ``torch~=2.4,<2.13.1`` in the
reference's pyproject.toml:71; 2.11.0 installed here).

Parity status: **pinned** — ``oracle/generate_golden.py`` runs the unmodified
reference (under ``oracle/mpi4py_shim``) on seeded inputs and commits its outputs
to ``tests/golden/``; ``tests/test_oracle_golden.py`` checks every function here
against those vectors and against the reference's own known-answer ``cdist`` tests
(tests/spatial/test_distances.py:14-188, 207-265).

Reference files restated (paths relative to /root/reference):
  heat/core/communication.py:197-254     chunk()            -> chunk
  heat/spatial/distance.py:47-64         _quadratic_expand  -> quadratic_expand
  heat/spatial/distance.py:32-44         _euclidian_fast    -> euclidian_fast
  heat/spatial/distance.py:17-29         _euclidian         -> euclidian
  heat/spatial/distance.py:136-156,366-414  cdist/_dist     -> cdist
  heat/core/statistics.py:162-190        local_argmin       -> argmin_rows
  heat/cluster/_kcluster.py:352-370      _assign_to_cluster -> assign_to_cluster
  heat/cluster/kmeans.py:76-103          _update_centroids  -> update_centroids
  heat/cluster/kmeans.py:105-148         fit                -> fit
  heat/cluster/_kcluster.py:398-415      predict            -> predict
  heat/core/rounding.py:126-164          clip (Q3: int64 -> float32 tensor)
  heat/core/types.py:915-959             promote_types (float32 * int64 -> float64)
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

INT64_MAX = 9223372036854775807


# --------------------------------------------------------------------------- sharding
def chunk(n: int, p: int, r: int) -> Tuple[int, int]:
    """(offset, rows) of rank ``r`` of ``p`` for ``n`` rows split along axis 0.

    heat/core/communication.py:236-245 — the first ``n % p`` ranks get one extra row.
    """
    c, rem = divmod(n, p)
    if rem > r:
        c += 1
        start = r * c
    else:
        start = r * c + rem
    return start, c


def shard(x: torch.Tensor, p: int) -> List[torch.Tensor]:
    """Row shards of ``x`` exactly as ``ht.array(x, split=0)`` slices them (factories.py:428-434)."""
    out = []
    for r in range(p):
        o, c = chunk(x.shape[0], p, r)
        out.append(x[o : o + c])
    return out


# --------------------------------------------------------------------------- distances
def quadratic_expand(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """|x-y|^2 = |x|^2 + |y|^2 - 2xy, clamped at 0 (heat/spatial/distance.py:59-64)."""
    x_norm = (x**2).sum(1).view(-1, 1)
    y_t = torch.transpose(y, 0, 1)
    y_norm = (y**2).sum(1).view(1, -1)
    dist = x_norm + y_norm - 2.0 * torch.mm(x, y_t)
    return torch.clamp(dist, 0.0, np.inf)


def euclidian_fast(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """sqrt of the quadratic expansion (heat/spatial/distance.py:44)."""
    return torch.sqrt(quadratic_expand(x, y))


def euclidian(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """torch.cdist (heat/spatial/distance.py:29)."""
    return torch.cdist(x, y)


def gaussian(x: torch.Tensor, y: torch.Tensor, sigma: float = 1.0) -> torch.Tensor:
    """exp(-cdist(x, y)^2 / (2 sigma^2)) (heat/spatial/distance.py:67-84)."""
    d2 = euclidian(x, y) ** 2
    return torch.exp(-d2 / (2 * sigma * sigma))


def gaussian_fast(x: torch.Tensor, y: torch.Tensor, sigma: float = 1.0) -> torch.Tensor:
    """Same through the quadratic expansion (heat/spatial/distance.py:87-101)."""
    d2 = quadratic_expand(x, y)
    return torch.exp(-d2 / (2 * sigma * sigma))


def manhattan(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """torch.cdist(p=1) (heat/spatial/distance.py:104-117)."""
    return torch.cdist(x, y, p=1)


def manhattan_fast(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """sum |x - y| by dimension expansion (heat/spatial/distance.py:120-133)."""
    return torch.sum(torch.abs(x.unsqueeze(1) - y.unsqueeze(0)), dim=2)


def pairwise(x: torch.Tensor, y: torch.Tensor, metric: str = "euclidean", expand: bool = False,
             sigma: float = 1.0) -> torch.Tensor:
    """One tile ``metric(X.larray, Y block)`` of ``_dist`` (heat/spatial/distance.py:262, 409-414, 441, 472) for
    cdist / rbf / manhattan (:136-206)."""
    if x.ndim != 2 or y.ndim != 2:
        raise NotImplementedError("Only 2D data matrices are currently supported")
    if x.shape[1] != y.shape[1]:
        raise ValueError("Inputs must have same shape[1]")
    a, b = _promote_pair(x, y)
    if metric == "euclidean":
        return euclidian_fast(a, b) if expand else euclidian(a, b)
    if metric == "gaussian":
        return gaussian_fast(a, b, sigma) if expand else gaussian(a, b, sigma)
    if metric == "manhattan":
        return manhattan_fast(a, b) if expand else manhattan(a, b)
    raise ValueError(metric)


def _promote_pair(x: torch.Tensor, y: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """heat/spatial/distance.py:392-403: common type, at least float32; fp64 if either is 64-bit float."""
    if x.dtype == torch.float64 or y.dtype == torch.float64:
        t = torch.float64
    elif x.dtype in (torch.int64,) or y.dtype in (torch.int64,):
        # heat promote_types(int64, float32) -> float64 (types.py lattice)
        t = torch.float64
    else:
        t = torch.float32
    return x.to(t), y.to(t)


def cdist(x_local: torch.Tensor, y: torch.Tensor, quadratic_expansion: bool = False) -> torch.Tensor:
    """Local block of ``ht.spatial.cdist(X, Y)`` for X.split in {0, None}, Y replicated.

    heat/spatial/distance.py:409-414: ``d.larray = metric(X.larray, Y.larray)``.
    """
    if x_local.ndim != 2 or y.ndim != 2:
        raise NotImplementedError("Only 2D data matrices are currently supported")
    if x_local.shape[1] != y.shape[1]:
        raise ValueError("Inputs must have same shape[1]")
    a, b = _promote_pair(x_local, y)
    return euclidian_fast(a, b) if quadratic_expansion else euclidian(a, b)


def argmin_rows(dist: torch.Tensor) -> torch.Tensor:
    """argmin(axis=1, keepdims=True): torch.min first-index semantics, int64 (N,1).

    heat/core/statistics.py:177 (``torch.min(dim=1, keepdim=True)``), indices round-trip
    through float64 (:188) and come back as int64 (_operations.py:520) — lossless here.
    """
    _, idx = torch.min(dist, dim=1, keepdim=True)
    return idx.double().to(torch.int64)


# --------------------------------------------------------------------------- Lloyd step
def assign_to_cluster(
    x_local: torch.Tensor, centers: torch.Tensor, eval_functional_value: bool = False
):
    """heat/cluster/_kcluster.py:352-370 with KMeans' metric (kmeans.py:67, quadratic_expansion=True)."""
    d = cdist(x_local, centers, quadratic_expansion=True)
    labels = argmin_rows(d)
    if not eval_functional_value:
        return labels
    # ht.norm(distances.min(axis=1), ord=2) ** 2  (linalg/basics.py:2191-2193: sum of squares -> sqrt; then **2)
    mins = torch.min(d, dim=1)[0]
    return labels, mins


def functional_value(mins_per_shard: Sequence[torch.Tensor]) -> torch.Tensor:
    """``ht.norm(min_j D, 2) ** 2`` with one SUM allreduce across shards (_kcluster.py:367-368)."""
    partial = [torch.sum(m * m) for m in mins_per_shard]  # vector_norm: |x|^2 summed locally
    tot = partial[0].clone()
    for p in partial[1:]:
        tot = tot + p
    return torch.sqrt(tot) ** 2


def update_centroids(
    x_shards: Sequence[torch.Tensor], labels_shards: Sequence[torch.Tensor], centers: torch.Tensor
) -> torch.Tensor:
    """heat/cluster/kmeans.py:88-103 — the k-long loop of full passes, faithfully.

    * ``selection`` is int64, ``x * selection`` promotes fp32 -> fp64 (types.py:938-939)       [Q1]
    * count = allreduce(sum) clipped with ``clip(1.0, iinfo.max)`` which returns a float32
      tensor (rounding.py:156-164) -> divisor is ``double(float32(count))``                 [Q3]
    * empty cluster: masked sum is 0, count clipped to 1 -> centroid moves to the origin     [Q2]
    * new[i] = allreduce(sum_rows(assigned / count)); cast to centers' dtype on setitem.
    """
    k = centers.shape[0]
    new = centers.clone()
    for i in range(k):
        cnt = torch.zeros((1, 1), dtype=torch.int64)
        for lab in labels_shards:
            cnt = cnt + (lab == i).to(torch.int64).sum(dim=0, keepdim=True)
        cnt_clipped = cnt.clamp(1.0, INT64_MAX)  # float32 tensor, like the reference
        row = None
        for xs, lab in zip(x_shards, labels_shards):
            sel = (lab == i).to(torch.int64)
            prom = torch.float64 if xs.dtype in (torch.float32, torch.float64) else torch.float64
            assigned = xs.to(prom) * sel.to(prom)
            part = (assigned / cnt_clipped.to(prom)).sum(dim=0, keepdim=True)
            row = part if row is None else row + part
        new[i : i + 1, :] = row.to(new.dtype)
    return new


def tol_as_compared(tol: float) -> float:
    """``inertia <= tol``: the scalar becomes ``ht.array(tol)`` = float32 before promotion
    (_operations.py:117-122; probe: fp64 inertia 9.9999999e-05 <= 1e-4 is False)."""
    return float(np.float32(tol))


@dataclass
class FitResult:
    cluster_centers: torch.Tensor  # (k, d), dtype of init
    labels: List[torch.Tensor]  # per shard, (n_r, 1) int64 — vs. pre-update centroids of last iter [Q5]
    n_iter: int
    inertia: torch.Tensor  # 0-dim, centers' dtype (sum of squared centroid shift) [Q6]
    inertia_history: List[float]


def fit(
    x_shards: Sequence[torch.Tensor],
    init: torch.Tensor,
    max_iter: int = 300,
    tol: Optional[float] = 1e-4,
) -> FitResult:
    """heat/cluster/kmeans.py:105-148 with the DNDarray-init branch (_kcluster.py:136-143)."""
    if init.ndim != 2:
        raise ValueError("passed centroids need to be two-dimensional")
    d = x_shards[0].shape[1]
    if init.shape[1] != d:
        raise ValueError("passed centroids do not match cluster count or data shape")
    centers = init.clone()
    n_iter = 0
    hist = []
    labels = None
    inertia = None
    for _ in range(max_iter):
        n_iter += 1
        labels = [assign_to_cluster(xs, centers) for xs in x_shards]
        new = update_centroids(x_shards, labels, centers)
        inertia = ((centers - new) ** 2).sum()
        hist.append(float(inertia))
        centers = new.clone()
        if tol is not None:
            t = torch.tensor(tol_as_compared(tol), dtype=torch.float32).to(
                torch.promote_types(inertia.dtype, torch.float32)
            )
            if bool(inertia.to(t.dtype) <= t):
                break
    return FitResult(centers, labels, n_iter, inertia, hist)


def predict(x_shards: Sequence[torch.Tensor], centers: torch.Tensor):
    """heat/cluster/_kcluster.py:398-415 — labels per shard + functional value."""
    labs, mins = [], []
    for xs in x_shards:
        l, m = assign_to_cluster(xs, centers, eval_functional_value=True)
        labs.append(l)
        mins.append(m)
    return labs, functional_value(mins)


# --------------------------------------------------------------------------- fast equivalent
def update_centroids_fast(
    x_shards: Sequence[torch.Tensor], labels_shards: Sequence[torch.Tensor], centers: torch.Tensor
) -> torch.Tensor:
    """Same result as :func:`update_centroids` up to fp64 summation order (<=1e-15 relative),
    in one pass (index_add in fp64).  Used by tests at sizes where the k-pass loop is too slow;
    validated against :func:`update_centroids` in tests/test_oracle_golden.py."""
    k, d = centers.shape
    sums = torch.zeros((k, d), dtype=torch.float64)
    cnt = torch.zeros((k,), dtype=torch.int64)
    for xs, lab in zip(x_shards, labels_shards):
        li = lab.view(-1)
        sums.index_add_(0, li, xs.to(torch.float64))
        cnt += torch.bincount(li, minlength=k)
    div = cnt.clamp(min=1).to(torch.float32).to(torch.float64)  # Q3
    return (sums / div.view(-1, 1)).to(centers.dtype)


# --------------------------------------------------------------------------- parity comparator
@dataclass
class LabelParity:
    n: int
    mismatches: int
    near_ties: int  # |d_ref - d_new| < 1e-6 * d  (north-star window)
    ref_rounding: int  # outside that window but inside the reference's own fp32 rounding window
    hard: int  # neither: a real disagreement


def compare_labels(
    x: torch.Tensor, centers_pre: torch.Tensor, ref_labels: torch.Tensor, new_labels: torch.Tensor
) -> LabelParity:
    """North-star label rule: labels match exactly except on near-ties |Δd| < 1e-6·d, judged in fp64.

    Rows that disagree outside that window are further classified: the reference evaluates
    ``fl(|x|² + |c|²) − 2·fl(x·c)`` in the data's precision, whose absolute error on d² is about
    ``(d+3)·eps·(|x|² + |c|² + 2|x||c|)`` — a disagreement inside *that* window is the reference's
    own rounding noise (its result changes with the BLAS summation order), not a defect of the
    implementation under test.  ``hard`` counts everything else and must be 0.
    """
    ref = ref_labels.view(-1).long()
    new = new_labels.view(-1).long()
    bad = (ref != new).nonzero().view(-1)
    res = LabelParity(int(ref.numel()), int(bad.numel()), 0, 0, 0)
    if bad.numel() == 0:
        return res
    xb = x[bad].double()
    c = centers_pre.double()
    ca, cb = c[ref[bad]], c[new[bad]]
    da = (xb - ca).norm(dim=1)
    db = (xb - cb).norm(dim=1)
    near = (da - db).abs() < 1e-6 * torch.maximum(da, db)
    eps = float(torch.finfo(x.dtype).eps) if x.dtype.is_floating_point else 2.0**-23
    feat = x.shape[1]
    xn = (xb * xb).sum(1)
    mag = lambda cc: xn + (cc * cc).sum(1) + 2 * xn.sqrt() * (cc * cc).sum(1).sqrt()
    win = (feat + 3) * eps * torch.maximum(mag(ca), mag(cb))
    rr = (~near) & ((da * da - db * db).abs() <= 2 * win)
    res.near_ties = int(near.sum())
    res.ref_rounding = int(rr.sum())
    res.hard = int(bad.numel()) - res.near_ties - res.ref_rounding
    return res


def centers_rel_err(ref: torch.Tensor, new: torch.Tensor) -> float:
    """max |ref - new| / max(|ref|) per the north-star ("within 1e-5 relative (fp32) or 1e-12 (fp64)"):
    element-wise relative error with the scale floored at the largest centroid coordinate times eps
    would blow up on coordinates that are ~0, so the scale is the row's infinity norm."""
    r, n = ref.double(), new.double()
    scale = r.abs().amax(dim=1, keepdim=True).clamp_min(1e-300)
    return float(((r - n).abs() / scale).max())
