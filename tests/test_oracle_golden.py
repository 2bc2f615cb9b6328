"""The oracle (oracle/kmeans_oracle.py) pinned against outputs of the unmodified reference
(tests/golden/, written by oracle/generate_golden.py) and the reference's own known-answer tests."""
import os

import numpy as np
import pytest
import torch

from cases import CASES, make_case
from helpers import assert_fit_matches, check_inputs, load_golden
from oracle import kmeans_oracle as orc

SLOW = {"q3_count_gt_2p24_f64"}


@pytest.mark.parametrize("name", [n for n in CASES])
def test_oracle_fit_reproduces_reference(name):
    spec = CASES[name]
    x, init = make_case(name)
    gold = load_golden(name)
    check_inputs(name, x, init, gold)
    shards = [x]
    res = orc.fit(shards, init, max_iter=spec["max_iter"], tol=spec["tol"])
    labels = torch.cat(res.labels)
    # same library, same call sequence, same thread count -> the oracle must be bit-identical
    assert res.n_iter == int(gold["n_iter"])
    assert torch.equal(res.cluster_centers, torch.from_numpy(gold["centers"]))
    assert np.array_equal(labels.view(-1).numpy(), gold["labels"].astype(np.int64))
    assert float(res.inertia) == float(gold["inertia"])
    if name not in SLOW:
        plabs, fv = orc.predict(shards, res.cluster_centers)
        assert np.array_equal(torch.cat(plabs).view(-1).numpy(), gold["predict_labels"].astype(np.int64))
        np.testing.assert_allclose(float(fv), float(gold["functional_value"]), rtol=1e-6)


@pytest.mark.parametrize("name", ["blobs_f32_d8_k6", "blobs_f64_d16_k8", "config1_spherical"])
def test_oracle_two_shards_matches_reference_np2(name):
    """np=2 sharding rule + rank-ordered sums (reference np=2 run is bit-identical to np=1 in fp32)."""
    spec = CASES[name]
    x, init = make_case(name)
    gold = load_golden(name)
    res = orc.fit(orc.shard(x, 2), init, max_iter=spec["max_iter"], tol=spec["tol"])
    assert_fit_matches(name, x, init, gold, res.cluster_centers, torch.cat(res.labels), res.n_iter,
                       res.inertia)


def test_fast_update_equals_faithful_loop():
    x, init = make_case("blobs_f32_d32_k64")
    labels = orc.assign_to_cluster(x, init)
    a = orc.update_centroids([x], [labels], init)
    b = orc.update_centroids_fast([x], [labels], init)
    assert orc.centers_rel_err(a, b) < 1e-6
    xd = x.double()
    a = orc.update_centroids([xd], [labels], init.double())
    b = orc.update_centroids_fast([xd], [labels], init.double())
    assert orc.centers_rel_err(a, b) < 1e-13


def test_chunk_rule():
    # heat/core/communication.py:236-245
    assert [orc.chunk(10, 3, r) for r in range(3)] == [(0, 4), (4, 3), (7, 3)]
    assert [orc.chunk(2, 4, r) for r in range(4)] == [(0, 1), (1, 1), (2, 0), (2, 0)]
    for n in (0, 1, 7, 100, 1001):
        for p in (1, 2, 3, 4, 8):
            parts = [orc.chunk(n, p, r) for r in range(p)]
            assert sum(c for _, c in parts) == n
            assert all(parts[i][0] + parts[i][1] == parts[i + 1][0] for i in range(p - 1))


def test_cdist_golden_and_known_answers():
    g = load_golden("cdist")
    for nm, dt in (("f32", torch.float32), ("f64", torch.float64)):
        X, Y = torch.from_numpy(g[f"X_{nm}"]), torch.from_numpy(g[f"Y_{nm}"])
        for q, tag in ((False, "direct"), (True, "quad")):
            d = orc.cdist(X, Y, quadratic_expansion=q)
            assert torch.equal(d, torch.from_numpy(g[f"D_{nm}_{tag}"]))
    # reference known-answer: ones vs zeros in 4-D -> 2.0 (tests/spatial/test_distances.py:14-40)
    d = orc.cdist(torch.ones(4, 4), torch.zeros(6, 4), quadratic_expansion=True)
    assert torch.equal(d, torch.full((4, 6), 2.0))
    assert np.array_equal(g["ones_zeros"], np.full((4, 6), 2.0, dtype=np.float32))
    # reference: comparison with torch.cdist on a ramp, atol 1e-5 (tests/spatial/test_distances.py:207-265)
    A = torch.arange(30, dtype=torch.float32).reshape(10, 3)
    B = torch.arange(30, 48, dtype=torch.float32).reshape(6, 3)
    assert torch.allclose(orc.cdist(A, B, True), torch.cdist(A, B), atol=1e-5)
    with pytest.raises(NotImplementedError):
        orc.cdist(torch.zeros(2, 2, 2), torch.zeros(2, 2))


def test_metric_goldens_rbf_manhattan_and_self_distances():
    """rbf / manhattan / cdist(X) restated by the oracle are bit-identical to the unmodified reference
    (tests/golden/metrics.npz from oracle/generate_golden.py --metrics; heat/spatial/distance.py:67-133, 237-258), and the
    reference's own rings at np=3 reproduce the rows of its one-process result (metrics_rings.json)."""
    import json

    from cases import METRIC_SIGMA

    g = load_golden("metrics")
    c = load_golden("cdist")
    for nm in ("f32", "f64"):
        X, Y = torch.from_numpy(c[f"X_{nm}"]), torch.from_numpy(c[f"Y_{nm}"])
        for q, tag in ((False, "direct"), (True, "quad")):
            assert torch.equal(orc.pairwise(X, Y, "gaussian", q, METRIC_SIGMA), torch.from_numpy(g[f"rbf_{nm}_{tag}"]))
            assert torch.equal(orc.pairwise(X, Y, "euclidean", q), torch.from_numpy(g[f"cdist_{nm}_{tag}"]))
            assert torch.equal(orc.pairwise(X, X, "euclidean", q), torch.from_numpy(g[f"cdist_self_{nm}_{tag}"]))
            assert torch.equal(orc.pairwise(X, X, "gaussian", q, METRIC_SIGMA),
                               torch.from_numpy(g[f"rbf_self_{nm}_{tag}"]))
        assert torch.equal(orc.pairwise(X, Y, "manhattan", False), torch.from_numpy(g[f"manhattan_{nm}_direct"]))
        assert torch.equal(orc.pairwise(X, Y, "manhattan", True), torch.from_numpy(g[f"manhattan_{nm}_expand"]))
        assert torch.equal(orc.pairwise(X, X, "manhattan", True), torch.from_numpy(g[f"manhattan_self_{nm}"]))
    # known answers of the reference's tests (tests/spatial/test_distances.py:42-75): ones vs zeros in 4-D
    assert torch.allclose(orc.pairwise(torch.ones(4, 4), torch.zeros(6, 4), "gaussian", True, 1.0),
                          torch.full((4, 6), float(np.exp(-2.0))))
    assert torch.equal(orc.pairwise(torch.ones(4, 4), torch.zeros(6, 4), "manhattan", True), torch.full((4, 6), 4.0))
    with open(os.path.join(os.path.dirname(__file__), "golden", "metrics_rings.json")) as f:
        rings = json.load(f)
    assert rings["np"] == 3 and max(rings["max_abs_diff_vs_np1"].values()) < 1e-5


def test_consumer_goldens_kmedians_kmedoids_knn():
    """The restated KMedians / KMedoids / kNN (oracle/consumers_oracle.py) against the unmodified reference
    (tests/golden/consumers.npz): labels, iteration counts and kNN classes identical, centroids to rounding."""
    from cases import consumer_inputs
    from oracle import consumers_oracle as con

    g = load_golden("consumers")
    inp = consumer_inputs()
    for key in ("x", "init", "x_test"):
        assert np.array_equal(inp[key].numpy(), g[key]), key
    for dt, nm, tol in ((torch.float32, "f32", 1e-6), (torch.float64, "f64", 1e-13)):
        x, init = inp["x"].to(dt), inp["init"].to(dt)
        c, lab, n_iter, inertia = con.kmedians_fit(x, init, 30, 1e-4)
        assert n_iter == int(g[f"kmedians_{nm}_n_iter"])
        assert np.array_equal(lab.numpy(), g[f"kmedians_{nm}_labels"])
        np.testing.assert_allclose(c.numpy(), g[f"kmedians_{nm}_centers"], rtol=tol, atol=tol)
        np.testing.assert_allclose(float(inertia), float(g[f"kmedians_{nm}_inertia"]), atol=1e-10)
        pl, mins = con.assign_l1(x, c)
        assert np.array_equal(pl.numpy(), g[f"kmedians_{nm}_predict"])
        np.testing.assert_allclose(float(mins.double().sum()), float(g[f"kmedians_{nm}_fv"]), rtol=1e-5)
        c, lab, n_iter = con.kmedoids_fit(x, init, 30)
        assert n_iter == int(g[f"kmedoids_{nm}_n_iter"])
        assert np.array_equal(lab.numpy(), g[f"kmedoids_{nm}_labels"])
        assert np.array_equal(c.numpy(), g[f"kmedoids_{nm}_centers"])  # medoids are data rows: exact
        cls = con.knn_predict(x, inp["y"], inp["x_test"].to(dt), 5)
        assert np.array_equal(cls.numpy(), g[f"knn_{nm}_classes"])
        for p, tag in ((2, "bpkmeans"), (1, "bpkmedians")):  # BatchParallelKMeans / KMedians, one process, random_state 5
            c, n_iter = con.batch_parallel_fit([x], p, 4, 30, 1e-4, 5)
            assert n_iter == int(g[f"{tag}_{nm}_n_iter"])
            np.testing.assert_allclose(c.numpy(), g[f"{tag}_{nm}_centers"], rtol=tol, atol=tol)
            lab, fv = con.batch_parallel_predict(x, c, p)
            assert np.array_equal(lab.numpy(), g[f"{tag}_{nm}_predict"])
            np.testing.assert_allclose(fv, float(g[f"{tag}_{nm}_fv"]), rtol=1e-5)


def test_argmin_first_index_and_quirks():
    d = torch.tensor([[1.0, 0.5, 0.5], [2.0, 2.0, 3.0]])
    assert orc.argmin_rows(d).view(-1).tolist() == [1, 0]  # first index wins (statistics.py:177)
    # Q2: empty cluster moves to the origin; Q3: count clipped through float32
    x = torch.tensor([[1.0, 1.0], [3.0, 3.0]])
    c = torch.tensor([[2.0, 2.0], [100.0, 100.0]])
    lab = orc.assign_to_cluster(x, c)
    new = orc.update_centroids([x], [lab], c)
    assert torch.equal(new, torch.tensor([[2.0, 2.0], [0.0, 0.0]]))
    assert orc.tol_as_compared(1e-4) == float(np.float32(1e-4))
