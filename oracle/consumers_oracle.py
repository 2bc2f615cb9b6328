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
