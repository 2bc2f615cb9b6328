"""Seeded parity cases shared by oracle/generate_golden.py (reference side) and the tests.

Every case is (X global, init centroids) built from deterministic generators so that the committed
golden files only need to hold the reference's OUTPUTS plus a hash of the inputs.
"""
from __future__ import annotations

import torch

from heat_b200.synthetic import blobs_shard, initial_centroids, true_centres

CASES = {
    # BASELINE.json configs[0]: KMeans k=4 on create_spherical_dataset(4 x 2500 pts, 3 feats, fp32), np=2
    "config1_spherical": dict(n=10000, d=3, k=4, dtype="f32", max_iter=300, tol=1e-4, np2=True, kind="spherical"),
    "blobs_f32_d8_k6": dict(n=20000, d=8, k=6, dtype="f32", max_iter=300, tol=1e-4, np2=True, offset=1.5,
                            init_noise=2.0),
    "blobs_f32_d32_k64": dict(n=50000, d=32, k=64, dtype="f32", max_iter=300, tol=1e-4, offset=0.8,
                              init_noise=1.0),
    "blobs_f64_d16_k8": dict(n=30000, d=16, k=8, dtype="f64", max_iter=300, tol=1e-4, np2=True, offset=1.0,
                             init_noise=1.5),
    "blobs_f32_fixed5": dict(n=30011, d=32, k=64, dtype="f32", max_iter=5, tol=None, offset=0.8, init_noise=1.0),
    "overlap_f32_d4_k16": dict(n=20000, d=4, k=16, dtype="f32", max_iter=40, tol=1e-4, offset=0.7),
    "overlap_f64_d16_k8": dict(n=20000, d=16, k=8, dtype="f64", max_iter=40, tol=1e-6, offset=0.5),
    "empty_cluster_f32": dict(n=5000, d=8, k=5, dtype="f32", max_iter=10, tol=1e-4, kind="empty"),
    "mixed_f64data_f32init": dict(n=8000, d=16, k=8, dtype="f64", init_dtype="f32", max_iter=50, tol=1e-4),
    # float32 data + float64 centroids: the reference promotes both operands of cdist to float64 (distance.py:392-395)
    "mixed_f32data_f64init": dict(n=8000, d=16, k=8, dtype="f32", init_dtype="f64", max_iter=50, tol=1e-4),
    "replicated_f32": dict(n=6000, d=8, k=6, dtype="f32", max_iter=50, tol=1e-4, split=None),
    "odd_d5_k3_f32": dict(n=7001, d=5, k=3, dtype="f32", max_iter=50, tol=1e-4),
    "wide_d128_k32_f32": dict(n=6000, d=128, k=32, dtype="f32", max_iter=30, tol=1e-4),
    # k above the 256-centroid limit of the fused tensor-core kernel: the large-k path (BASELINE configs[3] shape class)
    "bigk_f32_d64_k320": dict(n=40000, d=64, k=320, dtype="f32", max_iter=6, tol=None, offset=0.8, init_noise=0.5),
    "q3_count_gt_2p24_f64": dict(n=(1 << 24) + 5, d=1, k=1, dtype="f64", max_iter=1, tol=None, kind="ramp"),
}

_DT = {"f32": torch.float32, "f64": torch.float64}


def make_case(name: str):
    s = CASES[name]
    dt = _DT[s["dtype"]]
    n, d, k = s["n"], s["d"], s["k"]
    kind = s.get("kind", "blobs")
    if kind == "ramp":
        x = ((torch.arange(n, dtype=torch.float64) % 7) + 0.5).view(-1, 1).to(dt)
        init = torch.zeros((1, 1), dtype=dt)
        return x, init
    if kind == "spherical":
        # heat/utils/data/spherical.py:7-54 semantics (4 balls on the diagonal), deterministic generator here
        x, _ = blobs_shard(n, d, k, dtype=dt, offset=4.0, seed=11, shuffled=False)
        init = initial_centroids(k, d, 4.0, 11, dt)
        return x, init
    offset = s.get("offset", 4.0)
    x, _ = blobs_shard(n, d, k, dtype=dt, offset=offset, seed=3, shuffled=True)
    init = initial_centroids(k, d, offset, 3, torch.float64)
    if "init_noise" in s:
        g = torch.Generator().manual_seed(99)
        init = true_centres(k, d, offset, 3, torch.float64) + s["init_noise"] * torch.randn(
            k, d, generator=g, dtype=torch.float64)
    init = init.to(_DT[s.get("init_dtype", s["dtype"])])
    if kind == "empty":
        init = init.clone()
        init[-1] = 1000.0  # nobody is closest to this centroid -> empty cluster -> origin (Q2)
    return x, init


#: sigma of the rbf goldens (oracle/generate_golden.py:METRIC_SIGMA)
METRIC_SIGMA = 2.5


def consumer_inputs():
    """Inputs of tests/golden/consumers.npz (oracle/generate_golden.py:consumer_inputs, same generator calls)."""
    import torch

    g = torch.Generator().manual_seed(11)
    k, d, n = 4, 5, 1500
    cent = 1.2 * torch.randn(k, d, generator=g, dtype=torch.float64)
    lab = torch.arange(n) % k
    x = (cent[lab] + torch.randn(n, d, generator=g, dtype=torch.float64)).to(torch.float32)
    x[17] = 0.0
    init = (cent + 0.5 * torch.randn(k, d, generator=g, dtype=torch.float64)).to(torch.float32)
    xt = (cent[torch.arange(90) % k] + 1.5 * torch.randn(90, d, generator=g, dtype=torch.float64)).to(torch.float32)
    return {"x": x, "init": init, "y": lab.clone(), "x_test": xt}
