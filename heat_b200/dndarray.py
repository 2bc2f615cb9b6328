"""Minimal mirror of Heat's ``DNDarray`` — just what the k-means path consumes and produces.

The real class is /root/reference/heat/core/dndarray.py:40-89 (constructor :65-74): a process-local
``torch.Tensor`` (``larray``) plus global metadata (``gshape``, ``dtype``, ``split``, ``device``,
``comm``, ``balanced``).  When the real Heat is importable, ``heat_b200.integration`` plugs the same
kernels under Heat's own DNDarray instead; this class exists so that the path (and its tests and
benchmark) run on a box that has no Heat/mpi4py installed.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch

from .communication import IN_PLACE, Communication, get_comm, sanitize_comm


class DNDarray:
    def __init__(
        self,
        array: torch.Tensor,
        gshape: Tuple[int, ...],
        dtype: torch.dtype,
        split: Optional[int],
        device: torch.device,
        comm: Communication,
        balanced: Optional[bool] = True,
    ):
        self.__array = array
        self.__gshape = tuple(int(s) for s in gshape)
        self.__dtype = dtype
        self.__split = split
        self.__device = device
        self.__comm = comm
        self.__balanced = balanced

    # -- metadata (dndarray.py:91-330) ---------------------------------------------------------------
    @property
    def larray(self) -> torch.Tensor:
        return self.__array

    @property
    def gshape(self) -> Tuple[int, ...]:
        return self.__gshape

    shape = gshape

    @property
    def lshape(self) -> Tuple[int, ...]:
        return tuple(self.__array.shape)

    @property
    def ndim(self) -> int:
        return len(self.__gshape)

    @property
    def dtype(self) -> torch.dtype:
        return self.__dtype

    @property
    def split(self) -> Optional[int]:
        return self.__split

    @property
    def device(self) -> torch.device:
        return self.__device

    @property
    def comm(self) -> Communication:
        return self.__comm

    @property
    def balanced(self) -> Optional[bool]:
        return self.__balanced

    def is_distributed(self) -> bool:
        return self.__split is not None and self.__comm.is_distributed()

    # -- the few operations the path needs -------------------------------------------------------------
    def resplit(self, axis: Optional[int] = None) -> "DNDarray":
        """Out-of-place redistribution; only ``-> None`` (replicate) is needed here
        (``init.resplit(None)``, heat/cluster/_kcluster.py:143)."""
        if axis == self.__split:
            return self
        if axis is None:
            if self.__split != 0:
                raise NotImplementedError("resplit(None) is implemented for split=0 arrays only")
            full = self.__comm.Allgatherv_rows(self.__array) if self.__comm.is_distributed() else self.__array
            return DNDarray(full, self.__gshape, self.__dtype, None, self.__device, self.__comm, True)
        raise NotImplementedError("only resplit(None) is implemented on this path")

    def copy(self) -> "DNDarray":
        return DNDarray(self.__array.clone(), self.__gshape, self.__dtype, self.__split, self.__device,
                        self.__comm, self.__balanced)

    def astype(self, dtype: torch.dtype, copy: bool = True) -> "DNDarray":
        arr = self.__array.to(dtype, copy=copy)
        return DNDarray(arr, self.__gshape, dtype, self.__split, self.__device, self.__comm, self.__balanced)

    def numpy(self) -> np.ndarray:
        """Global array as numpy (gathers split=0 arrays)."""
        return self.resplit(None).larray.detach().cpu().numpy() if self.__split == 0 else \
            self.__array.detach().cpu().numpy()

    def item(self):
        return self.__array.item()

    def __float__(self):
        return float(self.__array.item())

    def __len__(self):
        return self.__gshape[0]

    def __repr__(self):
        return (f"DNDarray(gshape={self.__gshape}, lshape={self.lshape}, dtype={self.__dtype}, "
                f"split={self.__split}, device={self.__device})")


def array(obj, dtype: Optional[torch.dtype] = None, split: Optional[int] = None,
          is_split: Optional[int] = None, device=None, comm: Optional[Communication] = None) -> DNDarray:
    """``ht.array`` for this path (heat/core/factories.py:151).

    ``split=0``  : ``obj`` is the *global* array on every rank; each rank keeps its chunk (:428-434).
    ``is_split=0``: ``obj`` is this rank's local block; the global shape is the sum of the blocks.
    """
    if split is not None and is_split is not None:
        raise ValueError("split and is_split are mutually exclusive parameters")
    comm = sanitize_comm(comm)
    t = obj if isinstance(obj, torch.Tensor) else torch.as_tensor(np.asarray(obj))
    if dtype is not None:
        t = t.to(dtype)
    elif t.dtype == torch.float64 and not isinstance(obj, (torch.Tensor, np.ndarray)):
        t = t.to(torch.float32)  # Heat's default float is float32 (types.py)
    if device is not None:
        t = t.to(device)
    if is_split is not None and is_split not in (0, -t.ndim if t.ndim else 0):
        raise NotImplementedError("only is_split=0 is supported on this path")
    if split is not None:
        if split < 0:
            split += t.ndim
        if not 0 <= split < max(t.ndim, 1):
            raise ValueError(f"split axis {split} out of range for a {t.ndim}-dimensional array")
        # any axis can be split (heat/core/factories.py:428-434); the k-means path itself accepts only axis 0 and
        # raises NotImplementedError for the others, exactly like the reference (tests/cluster/test_kmeans.py:68-100)
        _, lshape, slices = comm.chunk(t.shape, split)
        local = t[slices]
        return DNDarray(local, tuple(t.shape), t.dtype, split, t.device, comm, True)
    if is_split is not None:
        n = torch.tensor([t.shape[0]], dtype=torch.int64)
        if comm.is_distributed():
            n = n.to(t.device if t.is_cuda else "cpu")
            comm.Allreduce(IN_PLACE, n)
        gshape = (int(n.item()),) + tuple(t.shape[1:])
        return DNDarray(t, gshape, t.dtype, 0, t.device, comm, None)
    return DNDarray(t, tuple(t.shape), t.dtype, None, t.device, comm, True)
