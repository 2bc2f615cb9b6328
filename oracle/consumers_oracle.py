"""CPU restatement of the other consumers of the assignment pattern.  TEST INFRASTRUCTURE ONLY: the
product (heat_b200/) never imports this module.  Pinned against outputs of the unmodified reference
(tests/golden/consumers.npz, written by oracle/generate_golden.py --consumers; checked in tests/test_oracle_golden.py).

Restated, with the reference lines each function follows (paths relative to /root/reference):
  assign_l1        _KCluster._assign_to_cluster with metric = manhattan(expand=True)   heat/cluster/_kcluster.py:352-370,
                                                                                      heat/spatial/distance.py:120-133
  cluster_medians  the median step of KMedians / KMedoids._update_centroids            heat/cluster/kmedians.py:70-101
                   (all-zero rows dropped :76-79; linear interpolation                 heat/core/statistics.py:1684-1728)
  kmedians_fit     KMedians.fit                                                        heat/cluster/kmedians.py:105-147
  kmedoids_fit     KMedoids.fit / _update_centroids                                    heat/cluster/kmedoids.py:57-156
  knn_predict      KNeighborsClassifier.fit / predict                 heat/classification/kneighborsclassifier.py:55-135
  plus_plus_rows, kmex, batch_parallel_fit / _predict
                   _initialize_plus_plus, _kmex, _BatchParallelKCluster.fit / predict
                                                                    heat/cluster/batchparallelclustering.py:23-86, 171-331
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch

from . import kmeans_oracle as orc


def assign_l1(x: torch.Tensor, centers: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """labels (N,1) int64 and the row minima of manhattan_fast(x, centers)."""
    dist = orc.manhattan_fast(*orc._promote_pair(x, centers))
    mins, idx = torch.min(dist, dim=1, keepdim=True)
    return idx, mins.view(-1)


def cluster_medians(x: torch.Tensor, labels: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """medians [k, d] (NaN rows for empty clusters) and the number of rows that entered each median."""
    lab = labels.view(-1)
    med = torch.full((k, x.shape[1]), float("nan"), dtype=x.dtype)
    counts = torch.zeros(k, dtype=torch.int64)
    for i in range(k):
        assigned = x * (lab == i).to(torch.int64).view(-1, 1)
        rows = assigned.abs().sum(dim=1) != 0
        clean = assigned[rows]
        counts[i] = clean.shape[0]
        if clean.shape[0] == 0:
            continue
        s, _ = torch.sort(clean, dim=0)
        pos = 0.5 * (clean.shape[0] - 1)
        lo, hi = int(np.floor(pos)), int(np.ceil(pos))
        med[i] = s[lo] + (s[hi] - s[lo]) * (pos - np.floor(pos))
    return med, counts


def kmedians_fit(x: torch.Tensor, init: torch.Tensor, max_iter: int, tol: Optional[float]):
    centers = init.clone()
    n_iter, inertia, lab = 0, None, None
    for _ in range(max_iter):
        n_iter += 1
        lab, _ = assign_l1(x, centers)
        med, counts = cluster_medians(x, lab, centers.shape[0])
        if bool((counts == 0).any()):
            raise RuntimeError("empty cluster: the reference draws a random row here (not reproducible)")
        new = med.to(centers.dtype)
        inertia = ((centers - new) ** 2).sum()
        centers = new
        if tol is not None and bool(inertia <= tol):
            break
    return centers, lab, n_iter, inertia


def kmedoids_fit(x: torch.Tensor, init: torch.Tensor, max_iter: int):
    centers = init.clone()
    n_iter, lab = 0, None
    for _ in range(max_iter):
        n_iter += 1
        lab, _ = assign_l1(x, centers)
        med, counts = cluster_medians(x, lab, centers.shape[0])
        if bool((counts == 0).any()):
            raise RuntimeError("empty cluster: the reference draws a random row here (not reproducible)")
        new = centers.clone()
        for i in range(centers.shape[0]):
            dist = orc.manhattan_fast(x, med[i : i + 1].to(x.dtype))
            idx = int(torch.min(dist, dim=0).indices.item())  # argmin over all rows, first index
            new[i] = x[idx]
        if torch.equal(centers, new):
            break
        centers = new
    return centers, lab, n_iter


def knn_predict(x_train: torch.Tensor, y: torch.Tensor, x_test: torch.Tensor, n_neighbors: int) -> torch.Tensor:
    """y: integer labels (one-hot encoded like kneighborsclassifier.py:38-53) or an (n, classes) matrix."""
    if y.ndim == 1:
        one_hot = torch.zeros((y.shape[0], int(y.max()) + 1), dtype=torch.float32)
        one_hot[torch.arange(y.shape[0]), y.long()] = 1
        y = one_hot
    dist = orc.euclidian(*orc._promote_pair(x_test, x_train))
    _, idx = torch.topk(dist, n_neighbors, dim=1, largest=False)
    votes = y[idx.flatten()].reshape(idx.shape + (y.shape[1],)).sum(dim=1)
    return torch.argmax(votes, dim=1)


def plus_plus_rows(x: torch.Tensor, k: int, p: int, gen: Optional[torch.Generator]) -> torch.Tensor:
    """++ seeding on one shard: uniform first row, then rows drawn proportionally to the distance to the nearest chosen one
    (same generator calls as the reference after ``torch.manual_seed``: one randint, k-1 multinomials)."""
    cap = 2**24 - 1
    if x.shape[0] > cap:
        x = x[torch.randint(0, x.shape[0], (cap,), generator=gen)]
    chosen = [int(torch.randint(0, x.shape[0], (1,), generator=gen))]
    while len(chosen) < k:
        nearest = torch.cdist(x, x[chosen], p=p).min(dim=1).values
        chosen.append(int(torch.multinomial(nearest, 1, generator=gen)))
    return x[chosen]


def kmex(x: torch.Tensor, p: int, k: int, init, max_iter: int, tol: float, seed: Optional[int]):
    """One-shard k-means / k-medians: torch.cdist(p) labels, mean | torch.median (lower) update, empty clusters keep their
    centre, stop when allclose(new, old, atol=tol)."""
    gen = torch.Generator().manual_seed(seed) if seed is not None else None
    centers = init.clone() if isinstance(init, torch.Tensor) else plus_plus_rows(x, k, p, gen).clone()
    done = 0
    for done in range(1, max_iter + 1):
        lab = torch.cdist(x, centers, p=p).argmin(dim=1)
        before = centers.clone()
        for j in range(k):
            rows = x[lab == j]
            if rows.shape[0]:
                centers[j] = rows.median(dim=0).values if p == 1 else rows.mean(dim=0)
        if torch.allclose(centers, before, atol=tol):
            break
    return centers, done


def batch_parallel_fit(shards, p: int, k: int, max_iter: int, tol: float, random_state: Optional[int],
                       n_procs_to_merge: Optional[int] = None):
    """All ranks of _BatchParallelKCluster.fit in one process: returns (centres, n_iter as seen by rank 0)."""
    size = len(shards)
    seeds = [None if random_state is None else random_state + r for r in range(size)]
    local = [kmex(shards[r], p, k, "++", max_iter, tol, seeds[r]) for r in range(size)]
    cents, iters = [c for c, _ in local], [i for _, i in local]
    merge = n_procs_to_merge if n_procs_to_merge is not None else size
    alive = list(range(size))
    while len(alive) > 1:
        for pos in range(0, len(alive), merge):
            group = alive[pos:pos + merge]
            if len(group) > 1:
                root = group[0]
                pooled = torch.cat([cents[r] for r in group], dim=0)
                cents[root], extra = kmex(pooled, p, k, "++", max_iter, tol, seeds[root])
                iters[root] += extra
        alive = alive[::merge]
    return cents[0], iters[0]


def batch_parallel_predict(x: torch.Tensor, centers: torch.Tensor, p: int):
    lab = torch.cdist(x, centers, p=p).argmin(dim=1)
    diff = x - centers[lab]
    fv = float(torch.norm(diff, p="fro") ** 2) if p == 2 else float(torch.norm(diff, p=p, dim=1).sum())
    return lab.view(-1, 1).to(torch.int32), fv
