"""``cdist`` / ``rbf`` / ``manhattan`` with the reference's signatures (heat/spatial/distance.py:136-207) on the CUDA path.

``_dist`` below keeps the reference's layout rules (distance.py:209-479): ``X.split`` in {None, 0}, ``Y`` absent, replicated or
``split=0``; the result is ``split=0`` whenever X is, ``split=1`` for a replicated X against a distributed Y.  Every tile
``metric(X_local, Y_block)`` is one ``hk_pairwise`` call that writes straight into its column range of the local result.
Where the reference moves blocks with blocking ``Send/Probe/Recv`` around a ring (:262-359, :431-473), the blocks here travel
GPU to GPU (``torch.distributed`` isend/irecv: NCCL over NVLink) and step i+1's transfer is in flight while step i's tile is
computed.  For ``Y is None`` the reference computes half of the tiles and sends the transposes back; here every rank computes
its full row block (the tiles are cheap next to the exchange of results the transposes would need).
"""
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


def _local(t: torch.Tensor, dtype: torch.dtype, device=None) -> torch.Tensor:
    t = t.to(device=device if device is not None else t.device, dtype=dtype)
    if t.shape[0] > 0 and t.stride(1) != 1:
        t = t.contiguous()
    return t


def _ring(eng, comm, xl: torch.Tensor, stationary: torch.Tensor, out: torch.Tensor, kw: dict) -> None:
    """out[:, cols(r)] = metric(xl, block of rank r) for every rank r; blocks arrive around the ring."""
    counts = comm.row_counts(stationary.shape[0])
    displ = [0]
    for c in counts:
        displ.append(displ[-1] + c)
    size, rank = comm.size, comm.rank
    pending = comm.ring_post(stationary, 1, counts) if size > 1 else None
    eng.pairwise(xl, stationary, out[:, displ[rank]:displ[rank + 1]], **kw)
    for step in range(1, size):
        moving, reqs, sender = pending
        comm.ring_wait(reqs)
        if step + 1 < size:
            pending = comm.ring_post(stationary, step + 1, counts)
        eng.pairwise(xl, moving, out[:, displ[sender]:displ[sender + 1]], **kw)


def _dist(X: DNDarray, Y: Optional[DNDarray], kw: dict) -> DNDarray:
    if not isinstance(X, DNDarray):
        raise TypeError(f"X must be a DNDarray, but was {type(X)}")
    if len(X.shape) > 2:
        raise NotImplementedError("Only 2D data matrices are currently supported")
    if X.split not in (None, 0):
        raise NotImplementedError(
            f"Input split was X.split = {X.split}. Splittings other than 0 or None currently not supported.")
    comm = X.comm
    if Y is None:
        # distance.py:237-361
        t = X.dtype if X.dtype in _FLOATS else _promote(X.dtype, torch.float32)
        xl = _local(X.larray, t)
        eng = _engine.get_engine(xl.device)
        K = X.shape[0]
        out = torch.empty((xl.shape[0], K), dtype=t, device=xl.device)
        if X.split is None or not comm.is_distributed():
            eng.pairwise(xl, xl, out, **kw)
        else:
            _ring(eng, comm, xl, xl, out, kw)
        return DNDarray(out, (K, K), t, X.split, xl.device, comm, X.balanced)

    if not isinstance(Y, DNDarray):
        raise TypeError(f"Y must be a DNDarray, but was {type(Y)}")
    if len(Y.shape) > 2:
        raise NotImplementedError(
            f"Only 2D data matrices are supported, but input shapes were X: {X.shape}, Y: {Y.shape}")
    if X.comm is not Y.comm and (X.comm.size != Y.comm.size):
        raise NotImplementedError("Differing communicators not supported")
    if Y.split not in (None, 0):
        raise NotImplementedError(
            f"Input splits were X.split = {X.split}, Y.split = {Y.split}. Splittings other than 0 or None currently not supported.")
    if X.shape[1] != Y.shape[1]:
        raise ValueError("Inputs must have same shape[1]")
    split = X.split if X.split == 0 else (1 if Y.split == 0 else None)  # distance.py:375-390
    t = _promote(X.dtype, Y.dtype)
    xl = _local(X.larray, t)
    yl = _local(Y.larray, t, device=xl.device)
    eng = _engine.get_engine(xl.device)
    if X.split == 0 and Y.split == 0 and comm.is_distributed():
        out = torch.empty((xl.shape[0], Y.shape[0]), dtype=t, device=xl.device)  # distance.py:416-473
        _ring(eng, comm, xl, yl, out, kw)
    else:
        # replicated Y (:413-414), replicated X against the local block of Y (split=1, :409-410), or one process
        out = torch.empty((xl.shape[0], yl.shape[0]), dtype=t, device=xl.device)
        eng.pairwise(xl, yl, out, **kw)
    return DNDarray(out, (X.shape[0], Y.shape[0]), t, split, xl.device, comm, X.balanced)


def cdist(X: DNDarray, Y: Optional[DNDarray] = None, quadratic_expansion: bool = False) -> DNDarray:
    """Pairwise Euclidean distances between the rows of ``X`` and ``Y`` (distance.py:136-156)."""
    return _dist(X, Y, {"metric": "euclidean", "expand": bool(quadratic_expansion)})


def rbf(X: DNDarray, Y: Optional[DNDarray] = None, sigma: float = 1.0, quadratic_expansion: bool = False) -> DNDarray:
    """Gaussian kernel exp(-|x-y|^2 / (2 sigma^2)) between the rows of ``X`` and ``Y`` (distance.py:159-182)."""
    return _dist(X, Y, {"metric": "gaussian", "expand": bool(quadratic_expansion), "sigma": float(sigma)})


def manhattan(X: DNDarray, Y: Optional[DNDarray] = None, expand: bool = False) -> DNDarray:
    """Pairwise L1 distances between the rows of ``X`` and ``Y`` (distance.py:185-206)."""
    return _dist(X, Y, {"metric": "manhattan", "expand": bool(expand)})
