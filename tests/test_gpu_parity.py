"""GPU parity: the CUDA path (through the host mirror and the C ABI) against the outputs of the unmodified
reference (tests/golden/) and against the oracle on seeded inputs.  Bars (BASELINE.json north_star):
same n_iter; centroids within 1e-5 relative (fp32) / 1e-12 (fp64); labels exact except near-ties."""
import ctypes

import numpy as np
import pytest
import torch

import heat_b200 as hb
from cases import CASES, make_case
from helpers import CENTER_TOL, assert_fit_matches, check_inputs, load_golden
from heat_b200 import _lib, engine
from oracle import kmeans_oracle as orc

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda", 0)


def _paths_for(dtype):
    return ["simt", "generic", "tc", "auto"] if dtype == torch.float32 else ["simt", "generic", "auto"]


def _fit(name, path, sync_every=8):
    spec = CASES[name]
    x, init = make_case(name)
    hx = hb.array(x.to(DEV), split=spec.get("split", 0))
    km = hb.cluster.KMeans(n_clusters=init.shape[0], init=hb.array(init.to(DEV)), max_iter=spec["max_iter"],
                           tol=spec["tol"])
    km.kernel_path = path
    km.sync_every = sync_every
    km.fit(hx)
    return x, init, hx, km


def _tc_ok(x, init):
    eng = engine.get_engine(DEV)
    try:
        lab = torch.empty((min(x.shape[0], 512), 1), dtype=torch.int64, device=DEV)
        eng.assign(x[:512].to(DEV).contiguous(), init.to(DEV).contiguous(), lab, path="tc")
        return True
    except _lib.HKError:
        return False


@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("path", ["simt", "generic", "tc", "auto"])
def test_fit_matches_reference(name, path):
    spec = CASES[name]
    x, init = make_case(name)
    if path == "tc":
        if x.dtype != torch.float32 or init.dtype != torch.float32 or not _tc_ok(x, init):
            pytest.skip("tensor-core path does not cover this shape/dtype")
    gold = load_golden(name)
    check_inputs(name, x, init, gold)
    x, init, hx, km = _fit(name, path)
    assert km.cluster_centers_.split is None and km.labels_.split == spec.get("split", 0)
    assert km.labels_.dtype == torch.int64 and km.labels_.shape == (x.shape[0], 1)
    assert km.inertia_.shape == () and isinstance(km.n_iter_, int)
    # labels_ are judged against the pre-update centroids of the last iteration (quirk Q5): recover them
    # with the oracle from the reference's run
    res = orc.fit([x], init, max_iter=max(int(gold["n_iter"]) - 1, 0), tol=None) if int(gold["n_iter"]) > 1 else None
    pre = res.cluster_centers if res is not None else init
    assert_fit_matches(name, x, init, gold, km.cluster_centers_.larray, km.labels_.larray, km.n_iter_,
                       float(km.inertia_), pre_centers=pre.to(x.dtype))
    # predict + functional value (reference: _kcluster.py:398-415)
    pred = km.predict(hx)
    par = orc.compare_labels(x, torch.from_numpy(gold["centers"]).to(x.dtype),
                             torch.from_numpy(gold["predict_labels"].astype(np.int64)), pred.larray.cpu())
    assert par.hard == 0, par
    rtol = 1e-4 if x.dtype == torch.float32 else 1e-10
    np.testing.assert_allclose(float(km.functional_value_), float(gold["functional_value"]), rtol=rtol)


@pytest.mark.parametrize("sync_every", [1, 3, 1000])
def test_n_iter_independent_of_sync_interval(sync_every):
    name = "overlap_f32_d4_k16"
    gold = load_golden(name)
    _, _, _, km = _fit(name, "auto", sync_every)
    assert km.n_iter_ == int(gold["n_iter"])


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("n,d,k", [(1, 1, 1), (5, 3, 2), (257, 7, 3), (1000, 32, 64), (4099, 16, 8), (777, 64, 5),
                                   (3001, 32, 7), (130, 16, 100),
                                   (3000, 128, 40), (2049, 200, 9), (600, 5, 300), (5000, 33, 1100),
                                   # fused tensor-core kernel beyond k=64: multi-chunk TMEM sweep, nbuf 2-5, NA=4 plans
                                   (3000, 32, 96), (2500, 32, 160), (4000, 32, 256), (2100, 64, 96), (2600, 64, 160),
                                   (1900, 64, 256), (1500, 128, 96)])
def test_single_step_against_oracle(dtype, n, d, k):
    """One accumulate+finalize against the oracle for awkward shapes (ragged tiles, odd d, large k)."""
    g = torch.Generator().manual_seed(n * 7 + d * 3 + k)
    x = torch.randn(n, d, generator=g, dtype=torch.float64).to(dtype)
    c = torch.randn(k, d, generator=g, dtype=torch.float64).to(dtype)
    eng = engine.get_engine(DEV)
    for path in _paths_for(dtype):
        xd, cd = x.to(DEV), c.to(DEV).clone()
        part = torch.empty(k * (d + 1), dtype=torch.float64, device=DEV)
        lab = torch.empty(n, dtype=torch.int32, device=DEV)
        try:
            eng.lloyd_accumulate(xd, cd, part, labels=lab, path=path)
        except _lib.HKError:
            if path == "tc":
                continue
            raise
        ref_lab = orc.assign_to_cluster(x, c).view(-1)
        par = orc.compare_labels(x, c, ref_lab, lab.cpu().long())
        assert par.hard == 0, (path, par)
        # rows whose reference label depends on the BLAS summation order: a handful at most on unstructured data
        assert par.ref_rounding <= max(2, n // 500), (path, par)
        p = part.cpu().view(k, d + 1)
        new_lab = lab.cpu().long()
        exp = torch.zeros(k, d + 1, dtype=torch.float64)
        exp[:, :d].index_add_(0, new_lab, x.double())
        exp[:, d] = torch.bincount(new_lab, minlength=k).double()
        assert torch.equal(p[:, d], exp[:, d]), path
        tol = 2e-6 if dtype == torch.float32 else 1e-13
        scale = x.double().abs().max() * exp[:, d].clamp(min=1).view(-1, 1)
        assert float(((p[:, :d] - exp[:, :d]).abs() / scale).max()) < tol, path
        # finalize: quirks Q1-Q3 + shift
        c_out = torch.empty_like(cd)
        shift = torch.zeros((), dtype=dtype, device=DEV)
        state = torch.zeros(4, dtype=torch.int32, device=DEV)
        eng.lloyd_finalize(part, cd, c_out, True, 1e30, shift, state)
        want = orc.update_centroids_fast([x], [new_lab.view(-1, 1)], c)
        assert orc.centers_rel_err(want, c_out.cpu()) <= CENTER_TOL[dtype]
        assert state.cpu().tolist()[:2] == [1, 1]
        np.testing.assert_allclose(float(shift), float(((c - want) ** 2).sum()), rtol=1e-4)


def test_empty_and_strided_and_unaligned_shards():
    eng = engine.get_engine(DEV)
    k, d = 6, 8
    g = torch.Generator().manual_seed(5)
    c = torch.randn(k, d, generator=g).to(DEV)
    part = torch.full((k * (d + 1),), 7.0, dtype=torch.float64, device=DEV)
    eng.lloyd_accumulate(torch.empty((0, d), device=DEV), c, part)  # zero-row shard (dndarray.py:301-305)
    assert float(part.abs().sum()) == 0.0
    x = torch.randn(3000, d + 5, generator=g)
    xs = x.to(DEV)[:, 2 : 2 + d]  # row stride d+5, base offset 8 bytes: not bulk-copy eligible
    lab = torch.empty(3000, dtype=torch.int64, device=DEV)
    for path in ("simt", "generic", "auto"):
        eng.lloyd_accumulate(xs, c, part, labels=lab, path=path)
        ref = orc.assign_to_cluster(x[:, 2 : 2 + d].contiguous(), c.cpu()).view(-1)
        par = orc.compare_labels(x[:, 2 : 2 + d].contiguous(), c.cpu(), ref, lab.cpu())
        assert par.hard == 0
        assert float(part.view(k, d + 1)[:, d].sum()) == 3000.0


def _nan_inf_matrix(n, d, g):
    x = torch.randn(n, d, generator=g)
    x[3, 1] = float("nan")
    x[n // 2] = float("nan")
    x[7, 0] = float("inf")
    x[n - 5, d - 1] = -float("inf")
    x[11] = 0.0
    return x


@pytest.mark.parametrize("path,n,d,k,dtype", [("tc", 1000, 32, 64, torch.float32), ("tc", 700, 64, 160, torch.float32),
                                              ("simt", 1000, 32, 64, torch.float32),
                                              ("auto", 5000, 64, 320, torch.float32),
                                              ("auto", 3000, 128, 1024, torch.float32),
                                              # fp64 tensor-core kernel (d=16, k<=16) and its 128-byte-row twin
                                              ("simt", 1000, 16, 8, torch.float64), ("simt", 900, 16, 13, torch.float64),
                                              ("row128", 1000, 16, 8, torch.float64)])
def test_nan_inf_rows_on_every_path(path, n, d, k, dtype):
    """NaN / Inf rows (and an all-zero row) through the tensor-core filter, the 128-byte-row kernel and the large-k
    path: labels follow torch.min over clamp(d^2, 0) exactly like the reference (statistics.py:177, quirk Q4)."""
    g = torch.Generator().manual_seed(17)
    x = _nan_inf_matrix(n, d, g).to(dtype)
    c = torch.randn(k, d, generator=g).to(dtype)
    ref = orc.assign_to_cluster(x, c).view(-1)
    eng = engine.get_engine(DEV)
    lab = torch.empty(n, dtype=torch.int64, device=DEV)
    eng.assign(x.to(DEV), c.to(DEV), lab, path=path)
    got = lab.cpu()
    bad = torch.isnan(x).any(1) | torch.isinf(x).any(1)
    assert got[bad].tolist() == ref[bad].tolist()
    par = orc.compare_labels(x[~bad], c, ref[~bad], got[~bad])
    assert par.hard == 0, par
    # a NaN centroid: every distance to it is NaN and wins every row from its index on (first NaN sticks)
    c2 = c.clone()
    c2[k // 3, 2] = float("nan")
    ref2 = orc.assign_to_cluster(x, c2).view(-1)
    eng.assign(x.to(DEV), c2.to(DEV), lab, path=path)
    assert lab.cpu().tolist() == ref2.tolist()


@pytest.mark.parametrize("kind", ["randn", "uncentred", "overlap"])
def test_filter_cold_path_matches_oracle(kind):
    """Inputs on which the TF32 filter leaves many rows undecided (unstructured, far from the origin, overlapping):
    the candidate-only refinement must give the labels of the exact formula; the counters say it was exercised."""
    from heat_b200.synthetic import dataset_init, dataset_shard

    n, d, k = 60_000, 32, 64
    x, _ = dataset_shard(kind, n, d, k)
    c = dataset_init(kind, k, d)
    eng = engine.get_engine(DEV)
    xd, cd = x.to(DEV), c.to(DEV)
    ws = eng.row_workspace(n)
    eng.stats()
    for it, row_ws in enumerate((ws, ws, None)):  # workspace filled, workspace read, per-row bounds
        lab = torch.empty(n, dtype=torch.int32, device=DEV)
        part = torch.empty(k * (d + 1), dtype=torch.float64, device=DEV)
        fv = torch.zeros(1, dtype=torch.float64, device=DEV)
        eng.lloyd_accumulate(xd, cd, part, labels=lab, path="tc", row_ws=row_ws)
        st = eng.stats()
        ref = orc.assign_to_cluster(x, c).view(-1)
        par = orc.compare_labels(x, c, ref, lab.cpu().long())
        assert par.hard == 0, (kind, it, par)
        assert par.ref_rounding <= n // 2000, (kind, it, par)
        assert st["passes"] == 1 and st["rows"] == n
        if kind != "overlap":
            assert st["undecided_rows"] > 0  # the cold path ran
        assert st["all_centroid_rows"] == 0
        lab2 = torch.empty(n, dtype=torch.int32, device=DEV)
        eng.assign(xd, cd, lab2, fv, path="tc", row_ws=row_ws)
        assert torch.equal(lab, lab2)
        want_fv = float((orc.cdist(x, c, True).min(dim=1).values.double() ** 2).sum())
        np.testing.assert_allclose(float(fv), want_fv, rtol=1e-5)
        eng.stats()
    # the exact-FMA twin gives the same labels bit for bit
    lab3 = torch.empty(n, dtype=torch.int32, device=DEV)
    eng.assign(xd, cd, lab3, path="simt")
    assert torch.equal(lab, lab3)


def test_bigk_multichunk_and_tail_against_oracle():
    """Large-k path over more than one 1M-row launch plus a tail below 1024 rows (hk_lloyd_bigk.cu: row_base > 0,
    the exact-cdist tail branch), and the BASELINE configs[3] shape d=128, k=1024 at 1.1M rows."""
    eng = engine.get_engine(DEV)
    for n, d, k in ((2 * (1 << 20) + 500, 64, 320), (1_100_000, 128, 1024)):
        g = torch.Generator().manual_seed(n % 1000 + k)
        c = (2.0 * torch.randn(k, d, generator=g))
        x = c[torch.randint(0, k, (n,), generator=g)] + torch.randn(n, d, generator=g)
        xd, cd = x.to(DEV), c.to(DEV)
        part = torch.empty(k * (d + 1), dtype=torch.float64, device=DEV)
        lab = torch.empty(n, dtype=torch.int32, device=DEV)
        eng.lloyd_accumulate(xd, cd, part, labels=lab)
        assert eng.last_variant().startswith("bigk<"), eng.last_variant()
        got = lab.cpu().long()
        hard = rr = 0
        for r0 in range(0, n, 200_000):
            xs = x[r0 : r0 + 200_000]
            ref = orc.assign_to_cluster(xs, c).view(-1)
            par = orc.compare_labels(xs, c, ref, got[r0 : r0 + 200_000])
            hard += par.hard
            rr += par.ref_rounding
        assert hard == 0 and rr <= n // 5000, (n, d, k, hard, rr)
        p = part.cpu().view(k, d + 1)
        exp = torch.zeros(k, d + 1, dtype=torch.float64)
        exp[:, :d].index_add_(0, got, x.double())
        exp[:, d] = torch.bincount(got, minlength=k).double()
        assert torch.equal(p[:, d], exp[:, d])
        scale = x.double().abs().max() * exp[:, d].clamp(min=1).view(-1, 1)
        assert float(((p[:, :d] - exp[:, :d]).abs() / scale).max()) < 2e-6
        del xd, x


def test_nan_rows_follow_torch_min_semantics():
    # torch.min: a NaN distance wins, first NaN index sticks (quirk Q4)
    x = torch.tensor([[0.0, 0.0], [float("nan"), 1.0], [1.0, 1.0]])
    c = torch.tensor([[5.0, 5.0], [0.0, 0.0], [1.0, 1.0]])
    ref = orc.assign_to_cluster(x, c).view(-1)
    eng = engine.get_engine(DEV)
    lab = torch.empty(3, dtype=torch.int64, device=DEV)
    eng.assign(x.to(DEV), c.to(DEV), lab, path="simt")
    assert lab.cpu().tolist() == ref.tolist()


@pytest.mark.parametrize("dt", ["f32", "f64"])
def test_cdist_matches_reference(dt):
    g = load_golden("cdist")
    X, Y = torch.from_numpy(g[f"X_{dt}"]), torch.from_numpy(g[f"Y_{dt}"])
    for q, tag in ((False, "direct"), (True, "quad")):
        d = hb.spatial.cdist(hb.array(X.to(DEV), split=0), hb.array(Y.to(DEV)), quadratic_expansion=q)
        assert d.split == 0 and d.shape == (X.shape[0], Y.shape[0]) and d.dtype == X.dtype
        ref = torch.from_numpy(g[f"D_{dt}_{tag}"])
        atol = 1e-5 if dt == "f32" else 1e-8  # the reference's own tolerances (test_distances.py:207-265)
        assert torch.allclose(d.larray.cpu(), ref, atol=atol, rtol=0), (dt, tag, (d.larray.cpu() - ref).abs().max())
    ones = hb.spatial.cdist(hb.array(torch.ones(4, 4, device=DEV), split=0), hb.array(torch.zeros(6, 4, device=DEV)),
                            quadratic_expansion=True)
    assert torch.equal(ones.larray.cpu(), torch.full((4, 6), 2.0))
    # int inputs are promoted to float32 (distance.py:392-403)
    A = torch.arange(30, dtype=torch.int32).reshape(10, 3)
    dd = hb.spatial.cdist(hb.array(A.to(DEV), split=0), hb.array(A.to(DEV)), quadratic_expansion=True)
    assert dd.dtype == torch.float32
    assert torch.allclose(dd.larray.cpu(), torch.cdist(A.float(), A.float()), atol=1e-3)


@pytest.mark.parametrize("m,n,f", [(5000, 300, 64), (3000, 4096, 32), (2100, 4096, 128), (1024, 128, 96)])
def test_cdist_large_vs_oracle(m, n, f):
    """tcgen05 3xTF32 cdist kernel: squared distances inside the bound of the scheme (the dropped x_lo.y_lo term,
    2^-20 |x||y| with the factor 2 of the expansion, plus fp32 rounding of the formula), distances at the reference's
    own tolerance scale (tests/spatial/test_distances.py:207-265 use atol 1e-5 on O(1) values)."""
    g = torch.Generator().manual_seed(3)
    X = torch.randn(m, f, generator=g)
    Y = torch.randn(n, f, generator=g)
    got = hb.spatial.cdist(hb.array(X.to(DEV), split=0), hb.array(Y.to(DEV)), quadratic_expansion=True)
    assert engine.get_engine(DEV).last_variant().startswith("cdist_tc"), engine.get_engine(DEV).last_variant()
    got = got.larray.cpu().double()
    xd, yd = X.double(), Y.double()
    d2 = (xd * xd).sum(1, keepdim=True) + (yd * yd).sum(1) - 2.0 * xd @ yd.T
    xn, yn = xd.norm(dim=1, keepdim=True), yd.norm(dim=1)
    bound = 2.0**-19 * xn * yn + (f + 3) * 2.0**-22 * (xn * xn + yn * yn)
    assert bool(((got * got - d2).abs() <= bound + 1e-30).all()), float(((got * got - d2).abs() / bound).max())
    want = d2.clamp(min=0).sqrt()
    # distances here are O(sqrt(2 f)): 1e-5 relative to that scale
    assert float((got - want).abs().max()) <= 1e-5 * float(want.max()), float((got - want).abs().max())


@pytest.mark.parametrize("dt", ["f32", "f64"])
def test_rbf_manhattan_and_self_distances_match_reference(dt):
    """rbf / manhattan / cdist(X) through hk_pairwise against the unmodified reference's outputs
    (tests/golden/metrics.npz; heat/spatial/distance.py:67-133, 159-258) at its own tolerances
    (tests/spatial/test_distances.py:42-188: atol 1e-5 fp32, 1e-8 fp64)."""
    from cases import METRIC_SIGMA

    g, c = load_golden("metrics"), load_golden("cdist")
    X, Y = torch.from_numpy(c[f"X_{dt}"]).to(DEV), torch.from_numpy(c[f"Y_{dt}"]).to(DEV)
    hx, hy = hb.array(X, split=0), hb.array(Y)
    atol = 1e-5 if dt == "f32" else 1e-8
    eng = engine.get_engine(DEV)
    l0 = eng.launch_count()

    def close(d, name, shape):
        assert d.split == 0 and d.shape == shape and d.dtype == X.dtype and d.larray.is_cuda
        got, ref, tol = d.larray.cpu(), torch.from_numpy(g[name]), atol
        if name.startswith("cdist_self"):
            # self distances: the diagonal is sqrt(rounding of |x|^2 + |x|^2 - 2 x.x) on both sides (torch.cdist also expands above 25 rows),
            # not 0 — compare squared distances at the rounding of |x|^2 + |y|^2
            got, ref, tol = got * got, ref * ref, (1e-4 if dt == "f32" else 1e-12)
        assert torch.allclose(got, ref, atol=tol, rtol=0), (name, float((got - ref).abs().max()))

    for q, tag in ((False, "direct"), (True, "quad")):
        close(hb.spatial.rbf(hx, hy, sigma=METRIC_SIGMA, quadratic_expansion=q), f"rbf_{dt}_{tag}", (96, 40))
        close(hb.spatial.cdist(hx, quadratic_expansion=q), f"cdist_self_{dt}_{tag}", (96, 96))
        close(hb.spatial.rbf(hx, sigma=METRIC_SIGMA, quadratic_expansion=q), f"rbf_self_{dt}_{tag}", (96, 96))
        close(hb.spatial.manhattan(hx, hy, expand=q), f"manhattan_{dt}_{'expand' if q else 'direct'}", (96, 40))
    close(hb.spatial.manhattan(hx, expand=True), f"manhattan_self_{dt}", (96, 96))
    assert eng.launch_count() > l0
    # the reference's known answers (tests/spatial/test_distances.py:42-75, 120-150): ones vs zeros in 4-D
    o, z = hb.array(torch.ones(4, 4, device=DEV, dtype=X.dtype), split=0), hb.array(torch.zeros(6, 4, device=DEV, dtype=X.dtype))
    assert torch.allclose(hb.spatial.rbf(o, z, sigma=1.0, quadratic_expansion=True).larray.cpu().double(),
                          torch.full((4, 6), float(np.exp(-2.0)), dtype=torch.float64), atol=atol)
    assert torch.equal(hb.spatial.manhattan(o, z, expand=True).larray.cpu(), torch.full((4, 6), 4.0, dtype=X.dtype))


@pytest.mark.parametrize("m,n,f,sigma", [(3000, 512, 64, 8.0), (2100, 4096, 32, 5.0), (1500, 256, 128, 12.0)])
def test_rbf_large_on_the_tensor_core_kernel_vs_oracle(m, n, f, sigma):
    """Gaussian epilogue of the tcgen05 kernel (2^(d2 * -log2e / 2 sigma^2) on the 3xTF32 squared distance):
    |delta| <= exp(-t) * delta(d2) / (2 sigma^2) + 2^-21, far inside the reference's atol 1e-5 on values in (0, 1]."""
    g = torch.Generator().manual_seed(5)
    X, Y = torch.randn(m, f, generator=g), torch.randn(n, f, generator=g)
    got = hb.spatial.rbf(hb.array(X.to(DEV), split=0), hb.array(Y.to(DEV)), sigma=sigma, quadratic_expansion=True)
    assert engine.get_engine(DEV).last_variant().startswith("cdist_tc"), engine.get_engine(DEV).last_variant()
    want = orc.pairwise(X.double(), Y.double(), "gaussian", True, sigma)
    err = float((got.larray.cpu().double() - want).abs().max())
    assert err <= 2e-6, err
    assert float(want.max()) > 0.05  # the comparison is not one of zeros


def test_pairwise_writes_column_blocks_in_place():
    """What the rings do with every block (heat_b200/spatial.py:_ring): the tile lands in a column range of the wider
    local result, row stride kept, neighbouring columns untouched — aligned range (tensor-core kernel) and odd one."""
    g = torch.Generator().manual_seed(6)
    X, Y = torch.randn(2048, 64, generator=g), torch.randn(384, 64, generator=g)
    eng = engine.get_engine(DEV)
    want = orc.pairwise(X, Y, "euclidean", True)
    for c0 in (128, 131):
        out = torch.full((2048, 1024), -7.0, device=DEV)
        eng.pairwise(X.to(DEV), Y.to(DEV), out[:, c0:c0 + 384], "euclidean", True)
        assert eng.last_variant().startswith("cdist_tc" if c0 % 4 == 0 else "cdist_simt"), eng.last_variant()
        o = out.cpu()
        assert torch.allclose(o[:, c0:c0 + 384], want, atol=2e-5, rtol=0)
        assert bool((o[:, :c0] == -7.0).all()) and bool((o[:, c0 + 384:] == -7.0).all())
    # manhattan, large, both dtypes: sequential accumulation vs torch's, f * eps relative
    for dt, rtol in ((torch.float32, 2e-6), (torch.float64, 1e-13)):
        A, B = X[:700].to(dt), Y.to(dt)
        out = torch.empty(700, 384, dtype=dt, device=DEV)
        eng.pairwise(A.to(DEV), B.to(DEV), out, "manhattan", True)
        assert torch.allclose(out.cpu(), orc.pairwise(A.double(), B.double(), "manhattan", True).to(dt), rtol=rtol * 64, atol=0)


@pytest.mark.parametrize("nm", ["f32", "f64"])
def test_kmedians_kmedoids_knn_match_reference(nm):
    """KMedians / KMedoids / KNeighborsClassifier on the device against the unmodified reference's outputs
    (tests/golden/consumers.npz; heat/cluster/kmedians.py, kmedoids.py, heat/classification/kneighborsclassifier.py)."""
    from cases import consumer_inputs
    from test_gloo_multirank import _check_consumers

    dt = torch.float32 if nm == "f32" else torch.float64
    inp = consumer_inputs()
    eng = engine.get_engine(DEV)
    l0 = eng.launch_count()
    hx = hb.array(inp["x"].to(dt).to(DEV), split=0)
    init = hb.array(inp["init"].to(dt).to(DEV))
    km = hb.cluster.KMedians(n_clusters=4, init=init, max_iter=30, tol=1e-4).fit(hx)
    assert eng.last_variant().startswith(("assign_l1", "select_hist")), eng.last_variant()
    pred = km.predict(hx)
    kd = hb.cluster.KMedoids(n_clusters=4, init=init, max_iter=30).fit(hx)
    knn = hb.classification.KNeighborsClassifier(n_neighbors=5)
    knn.fit(hx, hb.array(inp["y"].to(DEV), split=0))
    cls = knn.predict(hb.array(inp["x_test"].to(dt).to(DEV), split=0))
    assert eng.launch_count() > l0 + 20
    assert km.cluster_centers_.larray.is_cuda and km.labels_.shape == (1500, 1) and km.labels_.dtype == torch.int64
    _check_consumers({"kmedians_centers": km.cluster_centers_.larray, "kmedians_labels": km.labels_.larray,
                      "kmedians_n_iter": km.n_iter_, "kmedians_inertia": float(km.inertia_),
                      "kmedians_predict": pred.larray, "kmedians_fv": float(km.functional_value_),
                      "kmedoids_centers": kd.cluster_centers_.larray, "kmedoids_labels": kd.labels_.larray,
                      "kmedoids_n_iter": kd.n_iter_, "knn_classes": cls.larray}, nm)


@pytest.mark.parametrize("nm", ["f32", "f64"])
def test_batch_parallel_clusterers_match_reference(nm):
    """BatchParallelKMeans / KMedians on the device against the unmodified reference (same seed: the same ++ rows are
    drawn), heat/cluster/batchparallelclustering.py."""
    from cases import consumer_inputs
    from test_gloo_multirank import _check_batch_parallel

    dt = torch.float32 if nm == "f32" else torch.float64
    hx = hb.array(consumer_inputs()["x"].to(dt).to(DEV), split=0)
    eng = engine.get_engine(DEV)
    l0 = eng.launch_count()
    res = {}
    for cls, tag, ini in ((hb.cluster.BatchParallelKMeans, "bpkmeans", "k-means++"),
                          (hb.cluster.BatchParallelKMedians, "bpkmedians", "k-medians++")):
        bp = cls(n_clusters=4, init=ini, max_iter=30, tol=1e-4, random_state=5).fit(hx)
        lab = bp.predict(hx)
        assert bp.cluster_centers_.larray.is_cuda and lab.larray.is_cuda and isinstance(bp.functional_value_, float)
        res[tag] = {"centers": bp.cluster_centers_.larray, "n_iter": bp.n_iter_, "labels": lab.larray, "fv": bp.functional_value_,
                    "dtype": lab.dtype}
    assert eng.launch_count() > l0 + 20
    _check_batch_parallel(res, 1, None, nm)
    # KMeans(init="batchparallel") starts from the batch-parallel centres (_kcluster.py:249-275)
    km = hb.cluster.KMeans(n_clusters=4, init="batchparallel", random_state=5, max_iter=3).fit(hx)
    assert km.n_iter_ >= 1 and km.cluster_centers_.shape == (4, 5) and km.labels_.shape == (1500, 1)
    with pytest.raises(NotImplementedError):
        hb.cluster.BatchParallelKMeans(init="random")
    with pytest.raises(ValueError):
        hb.cluster.BatchParallelKMeans(n_clusters=4).fit(hb.array(torch.zeros(8, 2, device=DEV)))  # not split


def test_l1_assignment_medians_topk_semantics():
    """Device kernels of the N4 consumers against their CPU restatement on inputs with ties, NaN and all-zero rows."""
    from oracle import consumers_oracle as con

    eng = engine.get_engine(DEV)
    g = torch.Generator().manual_seed(12)
    for dt in (torch.float32, torch.float64):
        # assignment: ties -> first index, NaN distance counts as the minimum (torch.min), narrow label types
        c = torch.tensor([[0.0, 0.0, 0.0], [2.0, 0.0, 0.0], [2.0, 0.0, 0.0], [0.0, 5.0, 1.0]], dtype=dt)
        x = torch.tensor([[1.0, 0.0, 0.0], [2.0, 0.1, 0.0], [float("nan"), 0.0, 0.0], [0.0, 4.0, 1.0],
                          [float("inf"), 0.0, 0.0]], dtype=dt)
        ref, mins = con.assign_l1(x, c)
        for ldt in (torch.int64, torch.int32, torch.uint8):
            lab = torch.empty(x.shape[0], dtype=ldt, device=DEV)
            eng.assign_l1(x.to(DEV), c.to(DEV), lab)
            assert lab.cpu().long().tolist() == ref.view(-1).tolist(), (dt, ldt)
        xr = torch.randn(70001, 9, generator=g, dtype=torch.float64).to(dt)
        cr = torch.randn(13, 9, generator=g, dtype=torch.float64).to(dt)
        ref, mins = con.assign_l1(xr, cr)
        lab = torch.empty(xr.shape[0], dtype=torch.int64, device=DEV)
        fv = torch.zeros(1, dtype=torch.float64, device=DEV)
        eng.assign_l1(xr.to(DEV), cr.to(DEV), lab, fv)
        bad = (lab.cpu() != ref.view(-1)).nonzero().view(-1)
        d_all = orc.manhattan_fast(xr[bad].double(), cr.double())
        gap = d_all.gather(1, lab.cpu()[bad].view(-1, 1)).view(-1) - d_all.min(dim=1).values
        assert bad.numel() <= 3 and bool((gap <= 1e-5 * d_all.min(dim=1).values).all())  # summation-order near-ties only
        np.testing.assert_allclose(float(fv), float(mins.double().sum()), rtol=1e-6)
        # wide rows: float64 at d = 100 does not fit the shared-memory tile and takes the row-per-thread kernels
        xw = torch.randn(5001, 100, generator=g, dtype=torch.float64).to(dt)
        cw = torch.randn(6, 100, generator=g, dtype=torch.float64).to(dt)
        ref, _ = con.assign_l1(xw, cw)
        lab = torch.empty(xw.shape[0], dtype=torch.int64, device=DEV)
        eng.assign_l1(xw.to(DEV), cw.to(DEV), lab)
        assert int((lab.cpu() != ref.view(-1)).sum()) == 0
        bd, bi = eng.nearest_rows_l1(xw.to(DEV), cw.to(DEV), 0)
        assert bi.cpu().tolist() == torch.min(orc.manhattan_fast(xw, cw), dim=0).indices.tolist()
        # rows of 8 / 16 / 32 / 64 features take the register-resident variant of the tiled kernels
        for dd in (8, 16, 32, 64):
            xa = torch.randn(20011, dd, generator=g, dtype=torch.float64).to(dt)
            ca = torch.randn(7, dd, generator=g, dtype=torch.float64).to(dt)
            ref, mins = con.assign_l1(xa, ca)
            lab = torch.empty(xa.shape[0], dtype=torch.int32, device=DEV)
            fv = torch.zeros(1, dtype=torch.float64, device=DEV)
            eng.assign_l1(xa.to(DEV), ca.to(DEV), lab, fv)
            assert int((lab.cpu().long() != ref.view(-1)).sum()) <= 1, dd
            np.testing.assert_allclose(float(fv), float(mins.double().sum()), rtol=1e-6)
            bd, bi = eng.nearest_rows_l1(xa.to(DEV), ca.to(DEV), 5)
            want = torch.min(orc.manhattan_fast(xa, ca), dim=0)
            assert (bi.cpu() - 5).tolist() == want.indices.tolist(), dd
            np.testing.assert_allclose(bd.cpu().numpy(), want.values.double().numpy(), rtol=1e-5)
        # medians: exact selection, zero rows dropped, even and odd cluster sizes, an empty cluster
        xm = torch.randn(200003, 7, generator=g, dtype=torch.float64).to(dt)
        xm[::1000] = 0.0
        xm[5::97, 3] = xm[6, 3]  # many equal values in one feature
        lm = torch.randint(0, 5, (xm.shape[0],), generator=g)
        lm[lm == 3] = 2  # cluster 3 stays empty
        med, counts = eng.cluster_medians(xm.to(DEV), lm.to(DEV), 5)
        rmed, rcounts = con.cluster_medians(xm, lm, 5)
        assert counts.cpu().tolist() == rcounts.tolist() and int(counts[3]) == 0
        ok = rcounts > 0
        np.testing.assert_allclose(med.cpu()[ok].numpy(), rmed[ok].numpy(), rtol=1e-6 if dt == torch.float32 else 1e-14, atol=0)
        # nearest rows (first index on ties) to k points
        pts = torch.cat([xm[1234:1235], xm[77:78] + 0.25, torch.zeros(1, 7, dtype=dt)])
        bd, bi = eng.nearest_rows_l1(xm.to(DEV), pts.to(DEV), 1000)
        dist = orc.manhattan_fast(xm, pts)
        want = torch.min(dist, dim=0)
        assert (bi.cpu() - 1000).tolist() == want.indices.tolist() and int(bi[2]) == 1000  # row 0 is the first zero row
        np.testing.assert_allclose(bd.cpu().numpy(), want.values.double().numpy(), rtol=1e-6)
        # top-k per row: ascending, lower index first on ties, NaN never before a number; class vote
        D = torch.randn(301, 157, generator=g, dtype=torch.float64).to(dt)
        D[:, 5] = D[:, 9]
        D[3, :10] = float("nan")
        D[7] = 1.0
        vals, idx = eng.topk_rows(D.to(DEV), 6)
        order = torch.argsort(torch.where(D != D, torch.full_like(D, float("inf")), D), dim=1, stable=True)[:, :6]
        assert torch.equal(idx.cpu(), order)
        assert torch.equal(vals.cpu(), D.gather(1, order))
        Y = torch.rand(157, 4, generator=g, dtype=torch.float64).to(dt)
        cls = eng.knn_vote(idx, Y.to(DEV))
        votes = Y[order.flatten()].reshape(301, 6, 4).sum(dim=1)
        assert (cls.cpu() != votes.argmax(dim=1)).sum() <= 1  # summation order of near-equal votes


def test_full_size_properties_config3():
    """BASELINE config 3 shard sizes through size-independent properties: counts sum to N, the k x d sums
    add up to the column sums of X (checksum of checksums), labels in range, repeatable bit-for-bit."""
    from heat_b200.synthetic import blobs_shard, initial_centroids

    n, d, k = 20_000_000, 32, 64
    x, _ = blobs_shard(n, d, k, device=DEV)
    c = initial_centroids(k, d).to(DEV)
    eng = engine.get_engine(DEV)
    part = torch.empty(k * (d + 1), dtype=torch.float64, device=DEV)
    lab = torch.empty(n, dtype=torch.uint8, device=DEV)
    eng.lloyd_accumulate(x, c, part, labels=lab)
    p = part.view(k, d + 1)
    assert float(p[:, d].sum()) == float(n)
    col = x.sum(dim=0, dtype=torch.float64)
    # fp32 partial sums are kept to <= 32 rows before they are widened: relative to sum|x| the error is ~1e-10
    mag = x.abs().sum(dim=0, dtype=torch.float64)
    assert float(((p[:, :d].sum(0) - col).abs() / mag).max()) < 1e-8
    assert int(lab.max()) < k
    cnt = torch.bincount(lab.long(), minlength=k).double()
    assert torch.equal(cnt, p[:, d])
    part2 = torch.empty_like(part)
    eng.lloyd_accumulate(x, c, part2)
    assert torch.equal(part, part2)  # fixed tile->CTA map + ordered reduction: deterministic
    # sample check against the oracle
    idx = torch.randint(0, n, (20000,), generator=torch.Generator().manual_seed(1))
    xs = x[idx.to(DEV)].cpu()
    ref = orc.assign_to_cluster(xs, c.cpu()).view(-1)
    par = orc.compare_labels(xs, c.cpu(), ref, lab[idx.to(DEV)].cpu().long())
    assert par.hard == 0, par


def test_lloyd_run_replays_graphs_and_equals_single_steps():
    """hk_lloyd_run (first step eager, the rest replayed from cached CUDA graphs) gives bit-identical centroids to the same
    number of hk_lloyd_step calls, counts every executed iteration, and respects the sticky convergence flag."""
    from heat_b200.synthetic import blobs_shard, initial_centroids

    n, d, k = 300_000, 32, 64
    x, _ = blobs_shard(n, d, k, device=DEV, offset=1.0)
    c0 = initial_centroids(k, d, 1.0).to(DEV)
    eng = engine.get_engine(DEV)
    res = []
    for mode in ("steps", "run", "run_nograph"):
        c, cp = c0.clone(), torch.empty_like(c0)
        sh = torch.zeros((), device=DEV)
        st = torch.zeros(4, dtype=torch.int32, device=DEV)
        ws = eng.row_workspace(n)
        g0 = eng.graph_launch_count()
        if mode == "steps":
            for _ in range(11):
                eng.lloyd_step(x, c, cp, False, 0.0, sh, st, False, row_ws=ws)
        else:
            eng.graph(mode == "run")
            eng.lloyd_run(x, c, cp, False, 0.0, sh, st, False, 11, row_ws=ws)
            eng.graph(True)
        torch.cuda.synchronize()
        assert st.cpu().tolist()[:2] == [0, 11]
        if mode == "run":
            assert eng.graph_launch_count() - g0 == 3  # 10 replayed steps = 1 x 8 + 2 x 1
        res.append((c.clone(), cp.clone(), float(sh)))
    for other in res[1:]:
        assert torch.equal(res[0][0], other[0]) and torch.equal(res[0][1], other[1]) and res[0][2] == other[2]
    # with a tolerance: iterations enqueued after convergence are no-ops and n_iter stays exact
    c, cp = c0.clone(), torch.empty_like(c0)
    sh = torch.zeros((), device=DEV)
    st = torch.zeros(4, dtype=torch.int32, device=DEV)
    eng.lloyd_run(x, c, cp, True, 1e-4, sh, st, False, 60, row_ws=eng.row_workspace(n))
    flag, it = st.cpu().tolist()[:2]
    ref = orc.fit([x.cpu()], c0.cpu(), max_iter=60, tol=1e-4)
    assert flag == 1 and it == ref.n_iter
    assert orc.centers_rel_err(ref.cluster_centers, c.cpu()) <= 1e-5


def test_row_workspace_is_owned_by_the_caller():
    """The |x| bound cache lives in a caller-owned buffer tied to the matrix content: filled by the first pass (flag word
    set), read afterwards; a zeroed buffer after the rows changed, or no buffer at all, give the same results bit for
    bit.  The library itself keeps no per-matrix state (two matrices at the same address cannot confuse it)."""
    eng = engine.get_engine(DEV)
    n, d, k = 50_000, 32, 64
    g = torch.Generator().manual_seed(9)
    x = torch.randn(n, d, generator=g).to(DEV)
    c = torch.randn(k, d, generator=g).to(DEV)
    ws = eng.row_workspace(n)
    assert ws.numel() == (n + 127) // 128 * 4 + 16 and int(ws.view(torch.int32)[-4]) == 0
    outs = []
    for row_ws in (ws, ws, None):
        part = torch.empty(k * (d + 1), dtype=torch.float64, device=DEV)
        lab = torch.empty(n, dtype=torch.int32, device=DEV)
        eng.lloyd_accumulate(x, c, part, labels=lab, path="tc", row_ws=row_ws)
        outs.append((part.clone(), lab.clone()))
    assert int(ws.view(torch.int32)[(n + 127) // 128]) == 1  # filled
    for o in outs[1:]:
        assert torch.equal(outs[0][0], o[0]) and torch.equal(outs[0][1], o[1])
    # the rows change in place (same address, same shape): the caller zeroes its workspace, results follow the new rows
    x.mul_(100.0)
    ws.zero_()
    part = torch.empty(k * (d + 1), dtype=torch.float64, device=DEV)
    lab = torch.empty(n, dtype=torch.int32, device=DEV)
    eng.lloyd_accumulate(x, c, part, labels=lab, path="tc", row_ws=ws)
    ref = orc.assign_to_cluster(x.cpu(), c.cpu()).view(-1)
    assert orc.compare_labels(x.cpu(), c.cpu(), ref, lab.cpu().long()).hard == 0
    with pytest.raises(_lib.HKError, match="row workspace"):
        eng.lloyd_accumulate(x, c, part, path="tc", row_ws=ws[:64])


def test_kmeanspp_init_on_device():
    """init="kmeans++" (k-means||, _kcluster.py:146-245) with hk_cdist / hk_assign doing the N-sized work: on
    well-separated blobs the fit finds every true centre; string inits keep the reference's error contract."""
    from heat_b200.synthetic import blobs_shard, true_centres

    n, d, k = 200_000, 32, 64
    x, _ = blobs_shard(n, d, k, device=DEV, offset=4.0, seed=33)
    km = hb.cluster.KMeans(n_clusters=k, init="kmeans++", max_iter=30, tol=1e-4, random_state=1)
    km.fit(hb.array(x, split=0))
    dist = torch.cdist(true_centres(k, d, 4.0, 33).double(), km.cluster_centers_.larray.cpu().double())
    # k-means|| + Lloyd may leave a few centres merged/split on 64 clusters; most are found exactly
    assert int((dist.min(dim=1).values < 0.2).sum()) >= k - 6
    assert km.cluster_centers_.shape == (k, d) and km.n_iter_ >= 1
    with pytest.raises(ValueError):
        hb.cluster.KMeans(n_clusters=k, init="kmeans++").fit(hb.array(x, split=0), oversampling=1)


@pytest.mark.parametrize("k,d", [(96, 32), (160, 32), (128, 32), (32, 64), (32, 32), (48, 64)])
def test_tc_many_tiles_per_cta(k, d):
    """The fused tensor-core kernel over tens of tiles per CTA for every TMEM plan (2, 4 and 8 accumulator buffers, 8-12
    stages): the barrier hand-offs between the roles are only exercised when the pipeline wraps many times.  Labels
    against the oracle, sums against an fp64 index_add over the kernel's own labels, twice (run-to-run bit equality)."""
    n = 600_000
    g = torch.Generator().manual_seed(k * 7 + d)
    cen = 1.5 * torch.randn(k, d, generator=g)
    x = cen[torch.randint(0, k, (n,), generator=g)] + torch.randn(n, d, generator=g)
    c = cen + 0.3 * torch.randn(k, d, generator=g)
    eng = engine.get_engine(DEV)
    xd, cd = x.to(DEV), c.to(DEV)
    outs = []
    for rep in range(2):
        part = torch.empty(k * (d + 1), dtype=torch.float64, device=DEV)
        lab = torch.empty(n, dtype=torch.int32, device=DEV)
        eng.lloyd_accumulate(xd, cd, part, labels=lab, path="tc", row_ws=eng.row_workspace(n) if rep else None)
        outs.append((part.clone(), lab.clone()))
    assert eng.last_variant().startswith("tc<"), eng.last_variant()
    assert torch.equal(outs[0][1], outs[1][1]) and torch.equal(outs[0][0], outs[1][0])
    got = outs[0][1].cpu().long()
    hard = rr = 0
    for r0 in range(0, n, 200_000):
        xs = x[r0 : r0 + 200_000]
        par = orc.compare_labels(xs, c, orc.assign_to_cluster(xs, c).view(-1), got[r0 : r0 + 200_000])
        hard += par.hard
        rr += par.ref_rounding
    assert hard == 0 and rr <= n // 5000, (k, d, hard, rr)
    p = outs[0][0].cpu().view(k, d + 1)
    exp = torch.zeros(k, d + 1, dtype=torch.float64)
    exp[:, :d].index_add_(0, got, x.double())
    exp[:, d] = torch.bincount(got, minlength=k).double()
    assert torch.equal(p[:, d], exp[:, d])
    scale = x.double().abs().max() * exp[:, d].clamp(min=1).view(-1, 1)
    assert float(((p[:, :d] - exp[:, :d]).abs() / scale).max()) < 2e-6
