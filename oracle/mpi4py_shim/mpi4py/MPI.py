"""TEST INFRASTRUCTURE ONLY — minimal `mpi4py.MPI` surface for the reference.

np == 1: pure Python, no communication (the reference short-circuits every
reduction when ``size == 1``, heat/core/communication.py:1061-1065).
np  > 1: start one Python process per rank with RANK / WORLD_SIZE / MASTER_ADDR /
MASTER_PORT set; buffer collectives are mapped onto torch.distributed (gloo).
Only what KMeans.fit(init=DNDarray) / predict / cdist(Y replicated) touch is
implemented (Allreduce, allreduce, allgather, bcast, Bcast, Barrier); the rest
raises NotImplementedError so silent mis-behaviour is impossible.
"""
from __future__ import annotations

import ctypes
import os
import pickle

import numpy as np
import torch

_RANK = int(os.environ.get("RANK", "0"))
_SIZE = int(os.environ.get("WORLD_SIZE", "1"))
_dist = None


def _ensure_dist():
    global _dist
    if _SIZE == 1:
        return None
    if _dist is None:
        import torch.distributed as dist

        if not dist.is_initialized():
            dist.init_process_group(
                "gloo",
                init_method=f"tcp://{os.environ.get('MASTER_ADDR', '127.0.0.1')}:{os.environ['MASTER_PORT']}",
                rank=_RANK,
                world_size=_SIZE,
            )
        _dist = dist
    return _dist


class Exception(RuntimeError):  # noqa: A001 - mirrors mpi4py.MPI.Exception
    pass


class Datatype:
    def __init__(self, name, np_dtype=None, torch_dtype=None):
        self.name = name
        self.np_dtype = np_dtype
        self.torch_dtype = torch_dtype
        self.is_predefined = True

    def Create_vector(self, *a, **k):
        raise NotImplementedError("derived datatypes are not supported by the mpi4py stand-in")

    def Commit(self):
        return self

    def Free(self):
        pass


BOOL = Datatype("BOOL", np.bool_, torch.bool)
UNSIGNED_CHAR = Datatype("UNSIGNED_CHAR", np.uint8, torch.uint8)
SIGNED_CHAR = Datatype("SIGNED_CHAR", np.int8, torch.int8)
SHORT = Datatype("SHORT", np.int16, torch.int16)
INT16_T = SHORT
INT = Datatype("INT", np.int32, torch.int32)
LONG = Datatype("LONG", np.int64, torch.int64)
FLOAT = Datatype("FLOAT", np.float32, torch.float32)
DOUBLE = Datatype("DOUBLE", np.float64, torch.float64)
COMPLEX = Datatype("COMPLEX", np.complex64, torch.complex64)
DOUBLE_COMPLEX = Datatype("DOUBLE_COMPLEX", np.complex128, torch.complex128)


class Op:
    _next = 1

    def __init__(self, name, fn=None):
        self.name = name
        self.fn = fn
        self.handle = Op._next
        Op._next += 1

    @classmethod
    def Create(cls, function, commute=False):
        return cls("USER", function)

    def Free(self):
        pass


SUM = Op("SUM")
PROD = Op("PROD")
MIN = Op("MIN")
MAX = Op("MAX")
LAND = Op("LAND")
LOR = Op("LOR")
LXOR = Op("LXOR")
BAND = Op("BAND")
BOR = Op("BOR")
BXOR = Op("BXOR")
MINLOC = Op("MINLOC")
MAXLOC = Op("MAXLOC")

IN_PLACE = object()
ANY_TAG = -1
ANY_SOURCE = -2
ORDER_C = 0
MODE_WRONLY = 1
MODE_CREATE = 2
MODE_RDONLY = 4


class memory:  # noqa: N801 - mirrors mpi4py.MPI.memory
    def __init__(self, address, nbytes):
        self.address = address
        self.nbytes = nbytes

    @classmethod
    def fromaddress(cls, address, nbytes, readonly=False):
        return cls(address, nbytes)


buffer = memory


class Status:
    def __init__(self):
        self.source = 0
        self.nbytes = 0

    def Get_source(self):
        return self.source

    def Get_count(self, datatype=None):
        # bytes without a datatype (mpi4py's default MPI.BYTE), elements of ``datatype`` otherwise
        if datatype is None or datatype.np_dtype is None:
            return self.nbytes
        return self.nbytes // np.dtype(datatype.np_dtype).itemsize


class Request:
    def Wait(self, status=None):
        return True

    wait = Wait


class File:
    @staticmethod
    def Open(*a, **k):
        raise NotImplementedError


def _as_tensor(buf):
    """(memory, count, Datatype) or numpy array -> torch tensor aliasing that memory."""
    if isinstance(buf, np.ndarray):
        return torch.from_numpy(buf.reshape(-1) if buf.ndim else buf.reshape(1))
    mem, count, dt = buf
    if isinstance(count, (tuple, list)):
        raise NotImplementedError("v-collectives are not supported by the mpi4py stand-in")
    n = int(count)
    itemsize = np.dtype(dt.np_dtype).itemsize
    raw = (ctypes.c_char * (n * itemsize)).from_address(mem.address)
    arr = np.frombuffer(raw, dtype=dt.np_dtype, count=n)
    return torch.from_numpy(arr)


def _ni(name):
    def f(self, *a, **k):
        raise NotImplementedError(f"mpi4py stand-in: Comm.{name} is not implemented")

    f.__doc__ = f"{name} (not implemented in the stand-in)"
    f.__name__ = name
    return f


class Comm:
    def __init__(self, rank=_RANK, size=_SIZE):
        self._rank = rank
        self._size = size

    # -- introspection -----------------------------------------------------
    def Get_rank(self):
        return self._rank

    def Get_size(self):
        return self._size

    rank = property(Get_rank)
    size = property(Get_size)

    def Dup(self):
        return type(self)(self._rank, self._size)

    def Free(self):
        pass

    def Split(self, color=0, key=0):
        if self._size == 1:
            return self.Dup()  # one process: every split is the process itself
        raise NotImplementedError("mpi4py stand-in: Split is implemented for one process only")

    def Barrier(self):
        d = _ensure_dist() if self._size > 1 else None
        if d is not None:
            d.barrier()

    # -- buffer collectives --------------------------------------------------
    def Allreduce(self, sendbuf, recvbuf, op=SUM):
        """Allreduce(sendbuf, recvbuf, op=SUM)"""
        if self._size == 1:
            if sendbuf is not IN_PLACE:
                _as_tensor(recvbuf).copy_(_as_tensor(sendbuf))
            return
        d = _ensure_dist()
        r = _as_tensor(recvbuf)
        if sendbuf is not IN_PLACE:
            r.copy_(_as_tensor(sendbuf))
        rop = {
            "SUM": d.ReduceOp.SUM,
            "PROD": d.ReduceOp.PRODUCT,
            "MIN": d.ReduceOp.MIN,
            "MAX": d.ReduceOp.MAX,
        }.get(op.name)
        if rop is not None and r.dtype != torch.bool:
            d.all_reduce(r, op=rop)
            return
        if op.name in ("LAND", "LOR"):
            t = r.to(torch.int32)
            d.all_reduce(t, op=d.ReduceOp.MIN if op.name == "LAND" else d.ReduceOp.MAX)
            r.copy_(t.to(r.dtype))
            return
        if op.fn is not None:  # user op: gather everything, fold in rank order
            parts = [torch.empty_like(r) for _ in range(self._size)]
            d.all_gather(parts, r.clone())
            acc = parts[0].clone()
            for p in parts[1:]:
                a_np, b_np = p.numpy().copy(), acc.numpy()
                op.fn(memoryview(a_np.view(np.uint8)), memoryview(b_np.view(np.uint8)), None)
            r.copy_(acc)
            return
        raise NotImplementedError(f"mpi4py stand-in: Allreduce op {op.name}")

    def Bcast(self, buf, root=0):
        """Bcast(buf, root=0)"""
        if self._size == 1:
            return
        _ensure_dist().broadcast(_as_tensor(buf), src=root)

    # -- point-to-point for small host (numpy) messages ------------------------
    # (only what factories.array(is_split=...) needs: Isend / Probe / Recv of a shape vector)
    def _send_np(self, arr, dest):
        d = _ensure_dist()
        raw = torch.from_numpy(np.frombuffer(np.ascontiguousarray(arr).tobytes(), dtype=np.uint8).copy())
        d.send(torch.tensor([raw.numel()], dtype=torch.int64), dst=dest)
        d.send(raw, dst=dest)

    def _recv_np(self, source):
        pend = getattr(self, "_pending", {})
        if source in pend and pend[source]:
            return pend[source].pop(0)
        d = _ensure_dist()
        n = torch.zeros(1, dtype=torch.int64)
        d.recv(n, src=source)
        raw = torch.empty(int(n.item()), dtype=torch.uint8)
        d.recv(raw, src=source)
        return raw.numpy().tobytes()

    def Isend(self, buf, dest=0, tag=0):
        """Isend(buf, dest, tag=0)"""
        if not isinstance(buf, np.ndarray):
            buf = _as_tensor(buf).numpy()  # heat's (memory, count, Datatype) spec of a contiguous torch tensor
        self._send_np(buf, dest)
        return Request()

    Send = Isend

    def Probe(self, source=ANY_SOURCE, tag=ANY_TAG, status=None):
        """Probe(source, tag, status)"""
        msg = self._recv_np(source)
        if not hasattr(self, "_pending"):
            self._pending = {}
        self._pending.setdefault(source, []).insert(0, msg)
        if status is not None:
            status.source = source
            status.nbytes = len(msg)
        return True

    def Recv(self, buf, source=ANY_SOURCE, tag=ANY_TAG, status=None):
        """Recv(buf, source, tag, status)"""
        msg = self._recv_np(source)
        if not isinstance(buf, np.ndarray):
            t = _as_tensor(buf)
            t.copy_(torch.from_numpy(np.frombuffer(msg, dtype=t.numpy().dtype, count=t.numel()).copy()))
            return
        np.copyto(buf.reshape(-1), np.frombuffer(msg, dtype=buf.dtype, count=buf.size))

    # -- pickled-object collectives -------------------------------------------
    def allreduce(self, obj, op=SUM):
        if self._size == 1:
            return obj
        objs = self.allgather(obj)
        acc = objs[0]
        for o in objs[1:]:
            if op.name == "SUM":
                acc = acc + o
            elif op.name == "PROD":
                acc = acc * o
            elif op.name == "MIN":
                acc = min(acc, o)
            elif op.name == "MAX":
                acc = max(acc, o)
            elif op.name == "LAND":
                acc = acc and o
            elif op.name == "LOR":
                acc = acc or o
            else:
                raise NotImplementedError(op.name)
        return acc

    def allgather(self, obj):
        if self._size == 1:
            return [obj]
        out = [None] * self._size
        _ensure_dist().all_gather_object(out, pickle.loads(pickle.dumps(obj)))
        return out

    def bcast(self, obj, root=0):
        if self._size == 1:
            return obj
        box = [obj]
        _ensure_dist().broadcast_object_list(box, src=root)
        return box[0]


for _n in (
    "Irecv Ssend Issend Bsend Ibsend Rsend Irsend Ibcast Exscan Iexscan Scan Iscan "
    "Reduce Ireduce Iallreduce Allgather Iallgather Allgatherv Iallgatherv Alltoall Ialltoall "
    "Alltoallv Ialltoallv Alltoallw Ialltoallw Gather Igather Gatherv Igatherv Scatter Iscatter "
    "Scatterv Iscatterv Iprobe Sendrecv"
).split():
    setattr(Comm, _n, _ni(_n))


class Intracomm(Comm):
    pass


Communicator = Comm

COMM_WORLD = Intracomm(_RANK, _SIZE)
COMM_SELF = Intracomm(0, 1)
COMM_NULL = None


def Get_library_version():
    return "mpi4py stand-in (torch.distributed/gloo) — test infrastructure, not an MPI library"


def Is_initialized():
    return True


def Is_finalized():
    return False
