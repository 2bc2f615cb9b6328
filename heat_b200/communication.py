"""Communicator for the k-means path: one process per GPU, torch.distributed for the plumbing.

Mirrors the parts of ``heat.core.communication`` this path touches
(/root/reference/heat/core/communication.py): the ``Communication`` interface (:80-123 —
``is_distributed``, ``chunk``), ``rank``/``size`` and the call-site convention
``comm.Allreduce(MPI.IN_PLACE, tensor, MPI.SUM)`` (heat/core/_operations.py:510).  The reduction of
the per-iteration k x (d+1) partials does not go through here on the GPU path: ``hk_lloyd_step`` exchanges them
inside its finish kernel through peer-mapped GPU memory (csrc/hk_finalize.cu, csrc/hk_comm.cu); this class only
carries the bootstrap (NCCL id, cudaIpc handles) and small host-side collectives.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional, Tuple

import torch



class _InPlace:
    """Sentinel for ``comm.Allreduce(IN_PLACE, buf, SUM)`` (mpi4py's ``MPI.IN_PLACE``)."""

    def __repr__(self):
        return "IN_PLACE"


IN_PLACE = _InPlace()
SUM = "SUM"


class Communication:
    """Base interface (reference: communication.py:80-123)."""

    def is_distributed(self) -> bool:
        raise NotImplementedError()

    def chunk(self, shape, split, rank=None, w_size=None):
        raise NotImplementedError()


def chunk_rows(n: int, size: int, rank: int) -> Tuple[int, int]:
    """(offset, rows) of ``rank`` — the split=0 partition rule (communication.py:236-245)."""
    c, rem = divmod(int(n), int(size))
    if rem > rank:
        c += 1
        start = rank * c
    else:
        start = rank * c + rem
    return start, c


class ProcessGroupCommunication(Communication):
    """One rank per process; collectives ride on ``torch.distributed`` (nccl on GPUs, gloo on CPU)."""

    def __init__(self, group=None):
        import torch.distributed as dist

        self._dist = dist
        self.group = group
        if dist.is_available() and dist.is_initialized():
            self.rank = dist.get_rank(group)
            self.size = dist.get_world_size(group)
        else:
            self.rank, self.size = 0, 1

    # -- reference interface -----------------------------------------------------------------------
    def is_distributed(self) -> bool:
        return self.size > 1

    def chunk(self, shape, split, rank=None, w_size=None):
        """(offset, local_shape, slices) like MPICommunication.chunk (communication.py:197-254)."""
        shape = tuple(int(s) for s in shape)
        if split is None:
            return 0, shape, tuple(slice(0, e) for e in shape)
        if split < 0:
            split += len(shape)
        if not 0 <= split < len(shape):
            raise ValueError(f"split axis {split} out of range for shape {shape}")
        rank = self.rank if rank is None else rank
        w_size = self.size if w_size is None else w_size
        if not isinstance(rank, int) or not isinstance(w_size, int):
            raise TypeError("rank and size must be integers")
        start, c = chunk_rows(shape[split], w_size, rank)
        lshape = tuple(c if i == split else s for i, s in enumerate(shape))
        slices = tuple(slice(start, start + c) if i == split else slice(0, s) for i, s in enumerate(shape))
        return start, lshape, slices

    def Allreduce(self, sendbuf, recvbuf: torch.Tensor, op=SUM) -> None:
        """In-place sum over ranks (communication.py:1089-1110); np == 1 short-circuits (:1064)."""
        if op != SUM:
            raise NotImplementedError("only SUM is used on this path")
        if sendbuf is not IN_PLACE and not (isinstance(sendbuf, str) and sendbuf == "IN_PLACE"):
            recvbuf.copy_(sendbuf)
        if self.size > 1:
            self._dist.all_reduce(recvbuf, op=self._dist.ReduceOp.SUM, group=self.group)

    def Allgatherv_rows(self, local: torch.Tensor) -> torch.Tensor:
        """Concatenate row blocks of all ranks (used by ``resplit(None)`` of small arrays)."""
        if self.size == 1:
            return local
        counts = [None] * self.size
        self._dist.all_gather_object(counts, int(local.shape[0]), group=self.group)
        # all_gather needs equal block sizes: pad every block to the largest one
        cmax = max(counts)
        tail = tuple(local.shape[1:])
        mine = torch.zeros((cmax,) + tail, dtype=local.dtype, device=local.device)
        mine[: local.shape[0]] = local
        parts = [torch.empty((cmax,) + tail, dtype=local.dtype, device=local.device) for _ in counts]
        self._dist.all_gather(parts, mine, group=self.group)
        return torch.cat([p[:c] for p, c in zip(parts, counts)], dim=0)

    def row_counts(self, n_local: int) -> list:
        """Rows held by every rank, in rank order (``counts_displs_shape``, communication.py:256-281, but from the
        actual local shapes so unbalanced arrays work)."""
        if self.size == 1:
            return [int(n_local)]
        out = [None] * self.size
        self._dist.all_gather_object(out, int(n_local), group=self.group)
        return [int(c) for c in out]

    def ring_post(self, stationary: torch.Tensor, step: int, counts: list):
        """Start step ``step`` of the ring of ``_dist`` (distance.py:262-282, 431-462): this rank's block goes to rank
        ``(rank + step) % size`` while the block of rank ``(rank - step) % size`` arrives.  Returns ``(moving,
        requests, sender)``; the transfer runs beside whatever is launched until ``ring_wait``."""
        dist = self._dist
        receiver = (self.rank + step) % self.size
        sender = (self.rank - step) % self.size
        moving = torch.empty((counts[sender],) + tuple(stationary.shape[1:]), dtype=stationary.dtype,
                             device=stationary.device)
        ops = []
        # zero-row blocks are not sent: both sides know the counts
        if stationary.shape[0] > 0:
            ops.append(dist.P2POp(dist.isend, stationary, self._global_rank(receiver), group=self.group))
        if moving.shape[0] > 0:
            ops.append(dist.P2POp(dist.irecv, moving, self._global_rank(sender), group=self.group))
        reqs = dist.batch_isend_irecv(ops) if ops else []
        return moving, reqs, sender

    @staticmethod
    def ring_wait(reqs) -> None:
        for r in reqs:
            r.wait()

    def _global_rank(self, group_rank: int) -> int:
        if self.group is None:
            return group_rank
        return self._dist.get_global_rank(self.group, group_rank)

    def allgather_bytes(self, payload: bytes) -> list:
        """Every rank's ``payload`` in rank order (small host-side objects: IPC handles, row counts)."""
        if self.size == 1:
            return [payload]
        out = [None] * self.size
        self._dist.all_gather_object(out, payload, group=self.group)
        return out

    def bcast_bytes(self, payload: Optional[bytes], root: int = 0) -> bytes:
        if self.size == 1:
            return payload
        box = [payload]
        self._dist.broadcast_object_list(box, src=root, group=self.group)
        return box[0]

    def __repr__(self):
        return f"ProcessGroupCommunication(rank={self.rank}, size={self.size})"


_WORLD: Optional[ProcessGroupCommunication] = None


def init_from_env(backend: Optional[str] = None) -> ProcessGroupCommunication:
    """Join the torchrun rendezvous described by RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT."""
    import torch.distributed as dist

    global _WORLD
    ws = int(os.environ.get("WORLD_SIZE", "1"))
    if ws > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            # the reference binds rank r to cuda:(r % device_count) (heat/core/devices.py:116-120)
            torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")) % torch.cuda.device_count())
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend)
    _WORLD = ProcessGroupCommunication()
    return _WORLD


def get_comm() -> ProcessGroupCommunication:
    """Default communicator (reference: communication.get_comm, communication.py:2510)."""
    global _WORLD
    if _WORLD is None:
        _WORLD = ProcessGroupCommunication()
    return _WORLD


def use_comm(comm: Optional[Communication] = None) -> None:
    """Set the default communicator (reference: communication.use_comm, communication.py:2540-2550)."""
    global _WORLD
    _WORLD = sanitize_comm(comm)


def sanitize_comm(comm: Optional[Communication]) -> Communication:
    """reference: communication.sanitize_comm (communication.py:2519-2537)."""
    if comm is None:
        return get_comm()
    if isinstance(comm, Communication):
        return comm
    raise TypeError(f"Unknown communication, must be instance of {Communication}")
