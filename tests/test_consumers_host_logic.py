"""Host logic of the N3/N4 rows in one process, with the oracle-backed CheckerEngine standing in for the device
(tests/checker_engine.py): layouts and dtype promotion of ``_dist``, the empty-cluster failsafe of KMedians / KMedoids,
replicated inputs, soft labels in kNN, argument checks of the batch-parallel clusterers."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture()
def hb():
    import heat_b200
    from checker_engine import CheckerEngine
    from heat_b200 import engine

    saved = engine.get_engine
    engine.set_engine_factory(lambda dev: CheckerEngine(dev))
    yield heat_b200
    engine.get_engine = saved
    engine._ENGINES.clear()


def test_dist_layouts_promotion_and_errors(hb):
    from oracle import kmeans_oracle as orc

    g = torch.Generator().manual_seed(2)
    X, Y = torch.randn(12, 3, generator=g), torch.randn(5, 3, generator=g)
    for split in (None, 0):
        d = hb.spatial.cdist(hb.array(X, split=split), hb.array(Y))
        assert d.split == split and d.shape == (12, 5) and torch.equal(d.larray, orc.pairwise(X, Y, "euclidean", False))
        s = hb.spatial.cdist(hb.array(X, split=split), quadratic_expansion=True)
        assert s.split == split and s.shape == (12, 12) and torch.equal(s.larray, orc.pairwise(X, X, "euclidean", True))
    # one process: a "split" Y is the whole Y; X replicated against it gives split=1 (distance.py:375-390)
    d = hb.spatial.manhattan(hb.array(X), hb.array(Y, split=0), expand=False)
    assert d.split == 1 and torch.equal(d.larray, orc.pairwise(X, Y, "manhattan", False))
    r = hb.spatial.rbf(hb.array(X, split=0), hb.array(Y, split=0), sigma=0.7)
    assert r.split == 0 and torch.equal(r.larray, orc.pairwise(X, Y, "gaussian", False, 0.7))
    # integers are promoted to float32, int64 / float64 to float64 (distance.py:239-251, 392-403)
    A = torch.arange(12, dtype=torch.int32).reshape(4, 3)
    assert hb.spatial.cdist(hb.array(A, split=0)).dtype == torch.float32
    assert hb.spatial.rbf(hb.array(A, split=0), hb.array(Y.double())).dtype == torch.float64
    with pytest.raises(NotImplementedError):
        hb.spatial.manhattan(hb.array(torch.zeros(2, 2, 2)))
    with pytest.raises(NotImplementedError):
        hb.spatial.rbf(hb.array(X, split=1), hb.array(Y))
    with pytest.raises(NotImplementedError):
        hb.spatial.cdist(hb.array(X, split=0), hb.array(Y, split=1))
    with pytest.raises(ValueError):
        hb.spatial.cdist(hb.array(X), hb.array(torch.zeros(3, 4)))
    with pytest.raises(TypeError):
        hb.spatial.rbf(X, hb.array(Y))


def test_kmedians_kmedoids_failsafe_and_replicated_input(hb):
    from cases import consumer_inputs
    from helpers import load_golden

    inp, g = consumer_inputs(), load_golden("consumers")
    x, init = inp["x"], inp["init"].clone()
    # replicated input (split=None): same result as the split one, no collective
    km = hb.cluster.KMedians(n_clusters=4, init=hb.array(init), max_iter=30, tol=1e-4).fit(hb.array(x))
    assert km.n_iter_ == int(g["kmedians_f32_n_iter"]) and km.labels_.split is None
    np.testing.assert_allclose(km.cluster_centers_.larray.numpy(), g["kmedians_f32_centers"], rtol=1e-6, atol=1e-6)
    # a centroid far away from every row owns no point: the failsafe puts a data row there (kmedians.py:82-95)
    init[2] = 1e6
    for cls in (hb.cluster.KMedians, hb.cluster.KMedoids):
        est = cls(n_clusters=4, init=hb.array(init), max_iter=1, random_state=3)
        est.fit(hb.array(x, split=0))
        c = est.cluster_centers_.larray
        assert est.n_iter_ == 1 and bool(torch.isfinite(c).all()) and float(c.abs().max()) < 1e3
        assert bool((x == c[2]).all(dim=1).any()), "the replaced centroid is a row of the data"
    # aliases and argument errors
    assert hb.cluster.KMedians(init="kmedians++").init == "probability_based"
    assert hb.cluster.KMedoids(init="kmedoids++").init == "probability_based"
    with pytest.raises(ValueError):
        hb.cluster.KMedians(n_clusters=4, init=hb.array(init)).fit(x)


def test_knn_soft_labels_and_checks(hb):
    from oracle import consumers_oracle as con

    g = torch.Generator().manual_seed(4)
    xt, xq = torch.randn(60, 4, generator=g), torch.randn(9, 4, generator=g)
    soft = torch.rand(60, 3, generator=g)
    knn = hb.classification.KNeighborsClassifier(n_neighbors=4)
    knn.fit(hb.array(xt, split=0), hb.array(soft, split=0))
    assert knn.outputs_2d_ is True
    got = knn.predict(hb.array(xq, split=0))
    assert got.shape == (9,) and got.dtype == torch.int64 and got.split == 0
    assert torch.equal(got.larray, con.knn_predict(xt, soft, xq, 4))
    y = torch.randint(0, 5, (60,), generator=g)
    knn.fit(hb.array(xt, split=0), hb.array(y, split=0))
    assert knn.outputs_2d_ is False and knn.y.shape == (60, int(y.max()) + 1)
    assert torch.equal(knn.predict(hb.array(xq)).larray, con.knn_predict(xt, y, xq, 4))
    with pytest.raises(TypeError):
        knn.fit(xt, hb.array(y))
    with pytest.raises(ValueError):
        knn.fit(hb.array(xt), hb.array(y[:10]))
    with pytest.raises(ValueError):
        knn.fit(hb.array(torch.zeros(3)), hb.array(torch.zeros(3)))


def test_batch_parallel_argument_checks(hb):
    BP = hb.cluster.BatchParallelKMeans
    for bad, exc in (({"n_clusters": 2.0}, TypeError), ({"n_clusters": 0}, ValueError), ({"max_iter": 1.5}, TypeError),
                     ({"max_iter": 0}, ValueError), ({"tol": 1}, TypeError), ({"tol": -1.0}, ValueError),
                     ({"random_state": "x"}, TypeError), ({"n_procs_to_merge": 1}, ValueError),
                     ({"n_procs_to_merge": 2.0}, TypeError), ({"init": 3}, TypeError), ({"init": "nope"}, ValueError),
                     ({"init": "random"}, NotImplementedError)):
        with pytest.raises(exc):
            BP(**bad)
    with pytest.raises(NotImplementedError):
        hb.cluster.BatchParallelKMedians(init="random")
    bp = BP(n_clusters=2)
    with pytest.raises(TypeError):
        bp.fit(torch.zeros(4, 2))
    with pytest.raises(ValueError):
        bp.fit(hb.array(torch.zeros(4, 2)))
    with pytest.raises(ValueError):
        bp.fit(hb.array(torch.zeros(4, 2, 2), split=0))
    with pytest.raises(RuntimeError):
        bp.predict(hb.array(torch.zeros(4, 2), split=0))
    x = hb.array(torch.randn(40, 2, generator=torch.Generator().manual_seed(1)), split=0)
    bp = BP(n_clusters=2, random_state=1).fit(x)
    with pytest.raises(ValueError):
        bp.predict(hb.array(torch.zeros(4, 3), split=0))
    assert bp.predict(x).dtype == torch.int32 and isinstance(bp.functional_value_, float)
