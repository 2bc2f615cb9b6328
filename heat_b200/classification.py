"""``KNeighborsClassifier`` with the reference's API (heat/classification/kneighborsclassifier.py:8-135) on the CUDA path:
distances through ``heat_b200.spatial.cdist`` (any layout of ``_dist``), the ``n_neighbors`` smallest entries of every
row by ``hk_topk_rows``, the class vote by ``hk_knn_vote``."""
from __future__ import annotations

from typing import Callable, Optional

import torch

from . import engine as _engine
from . import spatial
from .dndarray import DNDarray


class KNeighborsClassifier:
    def __init__(self, n_neighbors: int = 5, effective_metric_: Optional[Callable] = None):
        self.n_neighbors = n_neighbors
        self.effective_metric_ = effective_metric_ if effective_metric_ is not None else spatial.cdist
        self.x = None
        self.y = None
        self.n_samples_fit_ = -1
        self.outputs_2d_ = True
        self.classes_ = None

    @staticmethod
    def one_hot_encoding(x: DNDarray) -> DNDarray:
        """Reference: kneighborsclassifier.py:38-53 — float32 one-hot rows, as many columns as the largest label + 1."""
        lab = x.larray.reshape(-1).long()
        top = int(lab.max().item()) if lab.numel() else -1
        if x.split is not None and x.comm.is_distributed():
            top = max(int.from_bytes(b, "little", signed=True)
                      for b in x.comm.allgather_bytes(int(top).to_bytes(8, "little", signed=True)))
        nc = top + 1
        one_hot = torch.zeros((lab.shape[0], nc), dtype=torch.float32, device=lab.device)
        if lab.numel():
            one_hot[torch.arange(lab.shape[0], device=lab.device), lab] = 1
        return DNDarray(one_hot, (x.shape[0], nc), torch.float32, x.split, lab.device, x.comm, x.balanced)

    def fit(self, x: DNDarray, y: DNDarray):
        """Reference: kneighborsclassifier.py:55-104."""
        if not isinstance(x, DNDarray) or not isinstance(y, DNDarray):
            raise TypeError(f"x and y must be DNDarrays but were {type(x)} {type(y)}")
        if len(x.shape) != 2:
            raise ValueError(f"x must be two-dimensional, but was {len(x.shape)}")
        self.x = x
        self.n_samples_fit_ = x.shape[0]
        if x.shape[0] != y.shape[0]:
            raise ValueError(f"Number of samples x and y samples mismatch, got {x.shape[0]}, {y.shape[0]}")
        if len(y.shape) == 1:
            self.y = self.one_hot_encoding(y)
            self.outputs_2d_ = False
        elif len(y.shape) == 2:
            self.y = y
            self.outputs_2d_ = True
        else:
            raise ValueError(f"y needs to be one- or two-dimensional, but was {len(y.shape)}")

    def predict(self, x: DNDarray) -> DNDarray:
        """Reference: kneighborsclassifier.py:106-135."""
        train = self.x
        if x.split is None and train.split == 0:
            # replicated queries against split training rows would give a column-split matrix (distance.py:375-390); the
            # neighbour search needs whole rows, so the training rows are gathered instead (the result is the same)
            train = train.resplit(None)
        distances = self.effective_metric_(x, train)
        if distances.split not in (None, 0):
            raise NotImplementedError("the neighbour search needs whole rows of the distance matrix (split 0 or None)")
        dl = distances.larray
        eng = _engine.get_engine(dl.device)
        _, idx = eng.topk_rows(dl, int(self.n_neighbors))
        y = self.y.resplit(None).larray if self.y.split is not None else self.y.larray
        y = y.to(device=dl.device)
        if y.dtype not in (torch.float32, torch.float64):
            y = y.to(torch.float32)
        classes = eng.knn_vote(idx, y.contiguous())
        self.classes_ = DNDarray(classes, (distances.shape[0],), torch.int64, distances.split, dl.device, x.comm, x.balanced)
        return self.classes_
