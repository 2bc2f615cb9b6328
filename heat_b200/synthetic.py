"""Synthetic "create_spherical_dataset-style" blobs, generalised from
/root/reference/heat/utils/data/spherical.py:7-54: unit-radius balls around centres spaced ``offset`` apart.

The global array is defined chunk by chunk (1 Mi rows, seeded ``seed + 1000 * chunk_index``) so that every
rank count produces the same global rows; each rank generates only the chunks overlapping its shard.
"""
from __future__ import annotations

from typing import Tuple

import torch

from .communication import chunk_rows

CHUNK = 1 << 20


def true_centres(k: int, d: int, offset: float = 4.0, seed: int = 1, dtype=torch.float32) -> torch.Tensor:
    """Cluster centres: reference-style diagonal (+-offset, +-2*offset) for k=4, d=3 (spherical.py:33-50);
    otherwise ``offset * randn(k, d)``."""
    if k == 4 and d == 3:
        o = offset
        t = torch.tensor([[o, o, o], [2 * o, 2 * o, 2 * o], [-o, -o, -o], [-2 * o, -2 * o, -2 * o]])
        return t.to(dtype)
    g = torch.Generator().manual_seed(seed)
    return (offset * torch.randn(k, d, generator=g, dtype=torch.float64)).to(dtype)


def initial_centroids(k: int, d: int, offset: float = 4.0, seed: int = 1, dtype=torch.float32) -> torch.Tensor:
    """``T + 0.5 * randn(k, d)`` (seed 2) — handed identically to the reference and to the new path."""
    g = torch.Generator().manual_seed(seed + 1)
    t = true_centres(k, d, offset, seed, torch.float64)
    return (t + 0.5 * torch.randn(k, d, generator=g, dtype=torch.float64)).to(dtype)


def _chunk(ci: int, rows: int, k: int, d: int, centres: torch.Tensor, radius: float, seed: int,
           device, dtype, shuffled: bool, n_global: int) -> torch.Tensor:
    g = torch.Generator(device=device).manual_seed(seed + 1000 * ci)
    base = ci * CHUNK
    idx = torch.arange(base, base + rows, device=device)
    if shuffled:
        lab = (idx * 2654435761 + 12345) % 1000003 % k  # fixed pseudo-random permutation of labels
    else:
        per = (n_global + k - 1) // k
        lab = torch.clamp(idx // per, max=k - 1)  # cluster-contiguous (spherical.py:52-53)
    u = torch.randn(rows, d, generator=g, device=device, dtype=torch.float32)
    u = u / u.norm(dim=1, keepdim=True).clamp_min(1e-20)
    r = torch.rand(rows, 1, generator=g, device=device, dtype=torch.float32).pow(1.0 / d)
    x = centres.to(device=device, dtype=torch.float32)[lab] + radius * r * u
    return x.to(dtype)


def blobs_shard(n_global: int, d: int, k: int, rank: int = 0, size: int = 1, device="cpu",
                dtype=torch.float32, offset: float = 4.0, radius: float = 1.0, seed: int = 1,
                shuffled: bool = True) -> Tuple[torch.Tensor, int]:
    """Rows ``[off, off+rows)`` of the global blob matrix for ``rank`` of ``size`` -> (tensor, off)."""
    off, rows = chunk_rows(n_global, size, rank)
    centres = true_centres(k, d, offset, seed, torch.float32)
    out = torch.empty((rows, d), dtype=dtype, device=device)
    pos = off
    while pos < off + rows:
        ci = pos // CHUNK
        c0 = ci * CHUNK
        crow = min(CHUNK, n_global - c0)
        blk = _chunk(ci, crow, k, d, centres, radius, seed, device, dtype, shuffled, n_global)
        lo, hi = pos - c0, min(crow, off + rows - c0)
        out[pos - off : pos - off + (hi - lo)] = blk[lo:hi]
        pos += hi - lo
    return out, off


# ---- other input distributions (bench.py --data): how the path behaves off well-separated blobs ---------------------
KINDS = ("blobs", "overlap", "randn", "uncentred")


def dataset_shard(kind: str, n_global: int, d: int, k: int, rank: int = 0, size: int = 1, device="cpu",
                  dtype=torch.float32, seed: int = 1) -> Tuple[torch.Tensor, int]:
    """Rows of ``rank`` of the global matrix of the given kind (same chunk-seeded scheme as :func:`blobs_shard`):

    * ``blobs``     unit balls around centres ``4 * randn`` (the benchmark's create_spherical_dataset-style input)
    * ``overlap``   the same with centres ``0.8 * randn``: heavily overlapping balls
    * ``randn``     unstructured standard normal rows (no cluster structure at all)
    * ``uncentred`` uniform [0, 255) rows, far from the origin (image-like, cf. the reference's cityscapes benchmark,
      benchmarks/2020/kmeans/config.json)
    """
    if kind == "blobs":
        return blobs_shard(n_global, d, k, rank, size, device, dtype, 4.0, 1.0, seed)
    if kind == "overlap":
        return blobs_shard(n_global, d, k, rank, size, device, dtype, 0.8, 1.0, seed)
    if kind not in KINDS:
        raise ValueError(f"unknown dataset kind {kind}")
    off, rows = chunk_rows(n_global, size, rank)
    out = torch.empty((rows, d), dtype=dtype, device=device)
    pos = off
    while pos < off + rows:
        ci = pos // CHUNK
        c0 = ci * CHUNK
        crow = min(CHUNK, n_global - c0)
        g = torch.Generator(device=device).manual_seed(seed + 77 + 1000 * ci)
        if kind == "randn":
            blk = torch.randn(crow, d, generator=g, device=device, dtype=torch.float32)
        else:
            blk = torch.rand(crow, d, generator=g, device=device, dtype=torch.float32) * 255.0
        lo, hi = pos - c0, min(crow, off + rows - c0)
        out[pos - off : pos - off + (hi - lo)] = blk[lo:hi].to(dtype)
        pos += hi - lo
    return out, off


def dataset_init(kind: str, k: int, d: int, dtype=torch.float32, seed: int = 1) -> torch.Tensor:
    """Initial centroids matching :func:`dataset_shard`."""
    if kind == "blobs":
        return initial_centroids(k, d, 4.0, seed, dtype)
    if kind == "overlap":
        return initial_centroids(k, d, 0.8, seed, dtype)
    g = torch.Generator().manual_seed(seed + 5)
    if kind == "randn":
        return torch.randn(k, d, generator=g, dtype=torch.float64).to(dtype)
    return (torch.rand(k, d, generator=g, dtype=torch.float64) * 255.0).to(dtype)
