"""CPU-side checks: the C-ABI library loads and exports what include/hkmeans.h declares, the host mirror
keeps the reference's argument and error contract, and nothing silently falls back to the CPU."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import heat_b200 as hb
from heat_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "hkmeans.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(hk_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 15
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"libhkmeans.so does not export {name}"
    assert declared == set(_lib.SIGNATURES), "ctypes prototypes and the header drifted apart"
    assert _lib.load().hk_version() >= 100


def test_no_gpu_calls_fail_loudly_not_silently():
    lib = _lib.load()
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = ctypes.c_void_p()
    rc = lib.hk_create(ctypes.byref(h), 0)
    assert rc != 0 and lib.hk_last_error()
    x = hb.array(torch.randn(10, 3), split=0)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        hb.cluster.KMeans(2, init=hb.array(torch.randn(2, 3))).fit(x)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        hb.spatial.cdist(x, x)


def test_chunk_matches_reference_rule():
    lib = _lib.load()
    for n, p in ((10, 3), (2, 4), (100000007, 8), (0, 2)):
        tot = 0
        for r in range(p):
            off, rows = ctypes.c_int64(), ctypes.c_int64()
            assert lib.hk_chunk(n, p, r, ctypes.byref(off), ctypes.byref(rows)) == 0
            assert (off.value, rows.value) == hb.communication.chunk_rows(n, p, r)
            assert off.value == tot
            tot += rows.value
        assert tot == n
    assert lib.hk_chunk(10, 0, 0, None, None) != 0


def test_row_workspace_size_and_null_handle_calls():
    """hk_row_ws_bytes is host arithmetic (one float per 128-row tile + flag words); calls on a null handle fail with a
    status and a message instead of crashing."""
    lib = _lib.load()
    assert lib.hk_row_ws_bytes(0) == 16
    assert lib.hk_row_ws_bytes(128) == 20 and lib.hk_row_ws_bytes(129) == 24
    assert lib.hk_row_ws_bytes(100_000_000) == (781250 + 4) * 4
    assert lib.hk_row_ws_bytes(-5) == 0
    out = (ctypes.c_int64 * 6)()
    assert lib.hk_stats_read(None, out) != 0 and b"hk_stats_read" in lib.hk_last_error()
    assert lib.hk_comm_mode(None) == 0 and lib.hk_graph_launch_count(None) == 0
    assert lib.hk_comm_peer_export(None, 16, None) != 0
    assert lib.hk_lloyd_run(None, None, 0, 1, 1, 0, None, None, 1, 0, 0.0, None, None, 0, None, 0, 0, 1, None) != 0


def test_in_place_sentinel_and_allgather_bytes():
    from heat_b200.communication import IN_PLACE, get_comm

    comm = get_comm()
    t = torch.arange(4.0)
    comm.Allreduce(IN_PLACE, t)  # np == 1: no-op, the buffer is not copied onto itself
    assert t.tolist() == [0.0, 1.0, 2.0, 3.0]
    comm.Allreduce("IN_PLACE", t)  # the string spelling of older call sites still means in place
    src = torch.ones(4)
    comm.Allreduce(src, t)
    assert t.tolist() == [1.0] * 4
    assert comm.allgather_bytes(b"abc") == [b"abc"]
    assert repr(IN_PLACE) == "IN_PLACE"


def test_estimator_protocol_and_params():
    # reference: tests/cluster/test_kmeans.py:18-34
    km = hb.cluster.KMeans()
    assert km.get_params() == {"init": "random", "max_iter": 300, "n_clusters": 8, "random_state": None,
                               "tol": 0.0001}
    km.set_params(n_clusters=3, tol=None)
    assert km.n_clusters == 3 and km.tol is None
    with pytest.raises(ValueError):
        km.set_params(bogus=1)
    assert hb.cluster.KMeans(init="kmeans++").init == "probability_based"
    for attr in ("cluster_centers_", "labels_", "inertia_", "n_iter_", "functional_value_"):
        assert getattr(km, attr) is None


def test_exception_contract():
    # reference: tests/cluster/test_kmeans.py:68-100, heat/cluster/kmeans.py:123-124, _kcluster.py:118-142,411-412
    x = hb.array(torch.randn(12, 4), split=0)
    with pytest.raises(TypeError):
        hb.cluster.KMeans(2).fit(torch.randn(12, 4))
    with pytest.raises(ValueError):
        hb.cluster.KMeans(2).predict(torch.randn(12, 4))
    with pytest.raises(ValueError):
        hb.cluster.KMeans(2, init="no-such-init").fit(x)
    with pytest.raises(ValueError):  # wrong centroid count
        hb.cluster.KMeans(3, init=hb.array(torch.randn(2, 4))).fit(x)
    with pytest.raises(ValueError):  # wrong feature count
        hb.cluster.KMeans(2, init=hb.array(torch.randn(2, 5))).fit(x)
    with pytest.raises(ValueError):  # not 2-D
        hb.cluster.KMeans(2, init=hb.array(torch.randn(2, 4, 1))).fit(x)
    with pytest.raises(ValueError):
        hb.cluster.KMeans(2, init=hb.array(torch.randn(2, 4))).fit(x, oversampling=1)
    with pytest.raises(ValueError):
        hb.cluster.KMeans(2, init=hb.array(torch.randn(2, 4))).fit(x, iter_multiplier=0)
    with pytest.raises(NotImplementedError):  # 3-D data
        hb.cluster.KMeans(2, init=hb.array(torch.randn(2, 4))).fit(hb.array(torch.randn(3, 4, 2), split=0))
    # cdist: tests/spatial/test_distances.py:190-205
    with pytest.raises(NotImplementedError):
        hb.spatial.cdist(hb.array(torch.randn(3, 4, 2)), hb.array(torch.randn(3, 4)))
    with pytest.raises(ValueError):
        hb.spatial.cdist(hb.array(torch.randn(3, 4)), hb.array(torch.randn(3, 5)))
    with pytest.raises(TypeError):
        hb.spatial.cdist(torch.randn(3, 4), hb.array(torch.randn(3, 4)))


def test_reference_exception_list_on_split1_data():
    # reference: tests/cluster/test_kmeans.py:68-100 (test_exceptions), statement by statement, with a split=1 matrix
    x = hb.array(torch.randn(150, 4), split=1)
    assert x.split == 1
    k = 3
    with pytest.raises(NotImplementedError):
        hb.cluster.KMeans(n_clusters=k).fit(x)
    with pytest.raises(ValueError):
        hb.cluster.KMeans(n_clusters=k).set_params(foo="bar")
    with pytest.raises(ValueError):
        hb.cluster.KMeans(n_clusters=k, init="random_number").fit(x)
    with pytest.raises(NotImplementedError):
        hb.cluster.KMeans(n_clusters=k, init="batchparallel").fit(x)
    with pytest.raises(ValueError):
        hb.cluster.KMeans(n_clusters=k, init=np.array([1, 2, 3])).fit(x)
    with pytest.raises(ValueError):
        hb.cluster.KMeans(n_clusters=k).fit(x, oversampling=-1)
    with pytest.raises(ValueError):
        hb.cluster.KMeans(n_clusters=k).fit(x, iter_multiplier=-1)
    with pytest.raises(ValueError):
        hb.cluster.KMeans(n_clusters=k, init=hb.array([1, 2])).fit(x)
    with pytest.raises(NotImplementedError):
        hb.cluster.KMeans(n_clusters=k, init="probability_based").fit(x)
    # cdist on other splittings: tests/spatial/test_distances.py:190-205
    with pytest.raises(NotImplementedError):
        hb.spatial.cdist(x, x)


def test_array_factory_split_semantics():
    g = torch.arange(40, dtype=torch.float32).reshape(10, 4)
    a = hb.array(g, split=0)
    assert a.split == 0 and a.shape == (10, 4) and a.lshape == (10, 4) and a.dtype == torch.float32
    b = hb.array(g, is_split=0)
    assert b.shape == (10, 4) and b.split == 0
    c = hb.array(g)
    assert c.split is None and c.resplit(None) is c
    assert torch.equal(a.resplit(None).larray, g)
    with pytest.raises(ValueError):
        hb.array(g, split=0, is_split=0)
    comm = hb.get_comm()
    assert comm.chunk((10, 4), 0, rank=1, w_size=3) == (4, (3, 4), (slice(4, 7), slice(0, 4)))
    assert comm.chunk((10, 4), None) == (0, (10, 4), (slice(0, 10), slice(0, 4)))
    with pytest.raises(TypeError):
        hb.communication.sanitize_comm("nope")


def test_synthetic_blobs_are_rank_invariant():
    from heat_b200.synthetic import blobs_shard

    full, _ = blobs_shard(2_500_000, 4, 5)
    parts = [blobs_shard(2_500_000, 4, 5, r, 3)[0] for r in range(3)]
    assert torch.equal(full, torch.cat(parts))
