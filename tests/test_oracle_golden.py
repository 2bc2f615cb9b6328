"""The oracle (oracle/kmeans_oracle.py) pinned against outputs of the unmodified reference
(tests/golden/, written by oracle/generate_golden.py) and the reference's own known-answer tests."""
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
