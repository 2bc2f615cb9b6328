"""``cdist`` with the reference's signature (heat/spatial/distance.py:136-156) on the CUDA path."""
from __future__ import annotations

from typing import Optional

import torch

from . import engine as _engine
from .dndarray import DNDarray

_FLOATS = (torch.float32, torch.float64)


def _promote(a: torch.dtype, b: torch.dtype) -> torch.dtype:
    # heat/spatial/distance.py:392-403 — common type, at least float32
    if a == torch.float64 or b == torch.float64 or a == torch.int64 or b == torch.int64:
        return torch.float64
    return torch.float32


def cdist(X: DNDarray, Y: Optional[DNDarray] = None, quadratic_expansion: bool = False) -> DNDarray:
    """Pairwise Euclidean distances between the rows of ``X`` and ``Y``.

    Supported layouts: ``X.split`` in {0, None} with ``Y`` replicated (the KMeans layout,
    distance.py:409-414).  ``Y is None`` and ``Y.split == 0`` use the reference's ring exchange
    (distance.py:237-361, 416-473), which is outside the accelerated path.
    """
    if not isinstance(X, DNDarray):
        raise TypeError(f"X must be a DNDarray, but was {type(X)}")
    if len(X.shape) > 2:
        raise NotImplementedError("Only 2D data matrices are currently supported")
    if Y is None:
        if X.split is not None and X.comm.is_distributed():
            raise NotImplementedError("cdist(X) with a distributed X needs the ring exchange (not on this path)")
        Y = X if X.split is None else X.resplit(None)
    if not isinstance(Y, DNDarray):
        raise TypeError(f"Y must be a DNDarray, but was {type(Y)}")
    if len(Y.shape) > 2:
        raise NotImplementedError("Only 2D data matrices are currently supported")
    if X.comm is not Y.comm and (X.comm.size != Y.comm.size):
        raise NotImplementedError("Differing communicators not supported")
    if X.split not in (None, 0):
        raise NotImplementedError("Splittings other than 0 or None currently not supported.")
    if Y.split is not None:
        if Y.split != 0:
            raise NotImplementedError("Splittings other than 0 or None currently not supported.")
        if Y.comm.is_distributed():
            raise NotImplementedError("cdist with Y.split=0 needs the ring exchange (not on this path)")
    if X.shape[1] != Y.shape[1]:
        raise ValueError("Inputs must have same shape[1]")

    t = _promote(X.dtype, Y.dtype)
    xl = X.larray.to(t)
    yl = Y.larray.to(device=xl.device, dtype=t)
    if xl.shape[0] > 0 and xl.stride(1) != 1:
        xl = xl.contiguous()
    if yl.shape[0] > 0 and yl.stride(1) != 1:
        yl = yl.contiguous()
    eng = _engine.get_engine(xl.device)
    out = torch.empty((xl.shape[0], yl.shape[0]), dtype=t, device=xl.device)
    eng.cdist(xl, yl, out, quadratic_expansion=bool(quadratic_expansion), sqrt=True)
    return DNDarray(out, (X.shape[0], Y.shape[0]), t, X.split, xl.device, X.comm, X.balanced)
