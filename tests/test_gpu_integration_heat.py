"""The Heat binding on CUDA: ``heat_b200.integration.install()`` patches the UNMODIFIED reference (installed in
``baseline/_ref``, imported under ``oracle/mpi4py_shim``) and real ``ht.cluster.KMeans(init=DNDarray).fit/predict`` and
``ht.spatial.cdist`` on CUDA DNDarrays run through libhkmeans.so; results are compared with the reference's own CPU
outputs (tests/golden).  Reference entry points: heat/cluster/kmeans.py:105-148, heat/cluster/_kcluster.py:352-415,
heat/spatial/distance.py:32-44."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")
SHIM = os.path.join(ROOT, "oracle", "mpi4py_shim")

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ht():
    if not os.path.isdir(os.path.join(REF, "heat")):
        pytest.skip("baseline/_ref (pip install --target of the reference) is not present")
    for q in (REF, SHIM):
        if q not in sys.path:
            sys.path.insert(0, q)
    import heat

    import heat_b200.integration as hki

    assert hki.install() is True
    yield heat
    hki.uninstall()


@pytest.mark.parametrize("name", ["blobs_f32_d32_k64", "blobs_f64_d16_k8", "config1_spherical", "empty_cluster_f32",
                                  "overlap_f32_d4_k16", "replicated_f32"])
def test_heat_kmeans_runs_on_the_native_path(ht, name):
    from cases import CASES, make_case
    from helpers import assert_fit_matches, load_golden
    from heat_b200 import engine
    from oracle import kmeans_oracle as orc

    spec = CASES[name]
    x, init = make_case(name)
    gold = load_golden(name)
    dev = torch.device("cuda", 0)
    eng = engine.get_engine(dev)
    l0 = eng.launch_count()
    hx = ht.array(x, split=spec.get("split", 0), device="gpu")
    hc = ht.array(init, device="gpu")
    assert hx.larray.is_cuda
    km = ht.cluster.KMeans(n_clusters=init.shape[0], init=hc, max_iter=spec["max_iter"], tol=spec["tol"])
    km.fit(hx)
    assert eng.launch_count() > l0, "the patched KMeans.fit did not launch a kernel of libhkmeans.so"
    assert isinstance(km.cluster_centers_, ht.DNDarray) and km.cluster_centers_.larray.is_cuda
    assert km.cluster_centers_.split is None and km.labels_.split == spec.get("split", 0)
    assert km.labels_.dtype == ht.int64 and tuple(km.labels_.shape) == (x.shape[0], 1)
    n_ref = int(gold["n_iter"])
    res = orc.fit([x], init, max_iter=max(n_ref - 1, 0), tol=None) if n_ref > 1 else None
    pre = res.cluster_centers if res is not None else init
    assert_fit_matches(name, x, init, gold, km.cluster_centers_.larray, km.labels_.larray, km.n_iter_,
                       float(km.inertia_.larray), pre_centers=pre.to(x.dtype))
    pred = km.predict(hx)
    par = orc.compare_labels(x, torch.from_numpy(gold["centers"]).to(x.dtype),
                             torch.from_numpy(gold["predict_labels"].astype(np.int64)), pred.larray.cpu())
    assert par.hard == 0, par
    rtol = 1e-4 if x.dtype == torch.float32 else 1e-10
    np.testing.assert_allclose(float(km.functional_value_.larray), float(gold["functional_value"]), rtol=rtol)


def test_heat_cdist_runs_on_the_native_path(ht):
    from helpers import load_golden
    from heat_b200 import engine

    g = load_golden("cdist")
    eng = engine.get_engine(torch.device("cuda", 0))
    for dt, atol in (("f32", 1e-5), ("f64", 1e-8)):
        X, Y = torch.from_numpy(g[f"X_{dt}"]), torch.from_numpy(g[f"Y_{dt}"])
        l0 = eng.launch_count()
        d = ht.spatial.cdist(ht.array(X, split=0, device="gpu"), ht.array(Y, device="gpu"), quadratic_expansion=True)
        assert eng.launch_count() > l0
        assert d.split == 0 and d.larray.is_cuda
        ref = torch.from_numpy(g[f"D_{dt}_quad"])
        assert torch.allclose(d.larray.cpu(), ref, atol=atol, rtol=0)


def test_heat_rbf_and_manhattan_run_on_the_native_path(ht):
    from cases import METRIC_SIGMA
    from helpers import load_golden
    from heat_b200 import engine

    g, c = load_golden("metrics"), load_golden("cdist")
    eng = engine.get_engine(torch.device("cuda", 0))
    for dt, atol in (("f32", 1e-5), ("f64", 1e-8)):
        X, Y = torch.from_numpy(c[f"X_{dt}"]), torch.from_numpy(c[f"Y_{dt}"])
        hx, hy = ht.array(X, split=0, device="gpu"), ht.array(Y, device="gpu")
        for fn, name in ((lambda: ht.spatial.rbf(hx, hy, sigma=METRIC_SIGMA, quadratic_expansion=True), f"rbf_{dt}_quad"),
                         (lambda: ht.spatial.rbf(hx, hy, sigma=METRIC_SIGMA), f"rbf_{dt}_direct"),
                         (lambda: ht.spatial.manhattan(hx, hy, expand=True), f"manhattan_{dt}_expand"),
                         (lambda: ht.spatial.manhattan(hx, hy), f"manhattan_{dt}_direct"),
                         (lambda: ht.spatial.cdist(hx), f"cdist_self_{dt}_direct"),
                         (lambda: ht.spatial.rbf(hx, sigma=METRIC_SIGMA, quadratic_expansion=True), f"rbf_self_{dt}_quad")):
            l0 = eng.launch_count()
            d = fn()
            assert eng.launch_count() > l0, name
            assert d.split == 0 and d.larray.is_cuda
            got, ref = d.larray.cpu(), torch.from_numpy(g[name])
            if name.startswith("cdist_self"):
                # self distances: the diagonal is sqrt(rounding of |x|^2 + |x|^2 - 2 x.x) on both sides (torch.cdist also expands above 25 rows),
                # not 0 — compare squared distances at the rounding of |x|^2 + |y|^2
                got, ref, tol = got * got, ref * ref, (1e-4 if dt == "f32" else 1e-12)
            else:
                tol = atol
            assert torch.allclose(got, ref, atol=tol, rtol=0), (name, float((got - ref).abs().max()))


def test_heat_kmedians_kmedoids_batchparallel_knn_run_on_the_native_path(ht):
    """Heat's own KMedians / KMedoids / BatchParallelKMeans / BatchParallelKMedians / KNeighborsClassifier objects on CUDA
    DNDarrays after install(): the kernels of libhkmeans.so run, results equal the unmodified reference's CPU goldens."""
    from cases import consumer_inputs
    from heat_b200 import engine
    from test_gloo_multirank import _check_batch_parallel, _check_consumers

    inp = consumer_inputs()
    eng = engine.get_engine(torch.device("cuda", 0))
    hx = ht.array(inp["x"], split=0, device="gpu")
    init = ht.array(inp["init"], device="gpu")
    l0 = eng.launch_count()
    km = ht.cluster.KMedians(n_clusters=4, init=init, max_iter=30, tol=1e-4)
    km.fit(hx)
    assert eng.last_variant().startswith(("assign_l1", "select_hist")), eng.last_variant()
    assert isinstance(km.cluster_centers_, ht.DNDarray) and km.cluster_centers_.larray.is_cuda
    pred = km.predict(hx)
    kd = ht.cluster.KMedoids(n_clusters=4, init=init, max_iter=30)
    kd.fit(hx)
    knn = ht.classification.kneighborsclassifier.KNeighborsClassifier(n_neighbors=5)
    knn.fit(hx, ht.array(inp["y"], split=0, device="gpu"))
    cls = knn.predict(ht.array(inp["x_test"], split=0, device="gpu"))
    assert eng.last_variant() == "topk_rows<f32>", eng.last_variant()
    assert eng.launch_count() > l0 + 30
    _check_consumers({"kmedians_centers": km.cluster_centers_.larray, "kmedians_labels": km.labels_.larray,
                      "kmedians_n_iter": km.n_iter_, "kmedians_inertia": float(km._inertia.item()),
                      "kmedians_predict": pred.larray, "kmedians_fv": float(km.functional_value_.item()),
                      "kmedoids_centers": kd.cluster_centers_.larray, "kmedoids_labels": kd.labels_.larray,
                      "kmedoids_n_iter": kd.n_iter_, "knn_classes": cls.larray})
    res = {}
    for klass, tag, ini in ((ht.cluster.BatchParallelKMeans, "bpkmeans", "k-means++"),
                            (ht.cluster.BatchParallelKMedians, "bpkmedians", "k-medians++")):
        l1 = eng.launch_count()
        bp = klass(n_clusters=4, init=ini, max_iter=30, tol=1e-4, random_state=5)
        bp.fit(hx)
        lab = bp.predict(hx)
        assert eng.launch_count() > l1 + 10 and lab.dtype == ht.int32
        res[tag] = {"centers": bp.cluster_centers_.larray, "n_iter": bp.n_iter_, "labels": lab.larray,
                    "fv": bp.functional_value_, "dtype": torch.int32}
    _check_batch_parallel(res, 1, None)
