"""Size-independent properties of the host logic and of the oracle (CPU, hypothesis)."""
import torch
from hypothesis import given, settings
from hypothesis import strategies as st

from heat_b200.communication import chunk_rows
from oracle import kmeans_oracle as orc


@settings(max_examples=300, deadline=None)
@given(n=st.integers(0, 10**10), p=st.integers(1, 64))
def test_chunk_rows_partitions_the_rows_like_the_reference(n, p):
    # reference rule (heat/core/communication.py:236-245): n // p rows, one more on the first n % p ranks, contiguous
    parts = [chunk_rows(n, p, r) for r in range(p)]
    assert sum(c for _, c in parts) == n
    off = 0
    for r, (start, c) in enumerate(parts):
        assert start == off and c == n // p + (1 if r < n % p else 0)
        off += c


@settings(max_examples=60, deadline=None)
@given(seed=st.integers(0, 2**31 - 1), n=st.integers(1, 300), d=st.integers(1, 9), k=st.integers(1, 7),
       f64=st.booleans())
def test_oracle_labels_are_the_nearest_centroid_up_to_near_ties(seed, n, d, k, f64):
    # _assign_to_cluster (heat/cluster/_kcluster.py:352-370) against a brute-force fp64 argmin
    g = torch.Generator().manual_seed(seed)
    dt = torch.float64 if f64 else torch.float32
    x = torch.randn(n, d, generator=g, dtype=torch.float64).to(dt)
    c = torch.randn(k, d, generator=g, dtype=torch.float64).to(dt)
    lab = orc.assign_to_cluster(x, c).view(-1)
    brute = torch.cdist(x.double(), c.double()).argmin(dim=1)
    par = orc.compare_labels(x, c, brute, lab)
    assert par.hard == 0, par


@settings(max_examples=40, deadline=None)
@given(seed=st.integers(0, 2**31 - 1), n=st.integers(2, 400), d=st.integers(1, 6), k=st.integers(1, 5),
       parts=st.integers(1, 4))
def test_oracle_update_is_independent_of_the_sharding(seed, n, d, k, parts):
    # Q7: fp64 sums rounded once -> the centroids do not depend on how the rows are split over ranks
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, d, generator=g)
    c = torch.randn(k, d, generator=g)
    lab = orc.assign_to_cluster(x, c)
    whole = orc.update_centroids([x], [lab], c)
    cuts = [chunk_rows(n, parts, r) for r in range(parts)]
    xs = [x[s:s + m] for s, m in cuts]
    ls = [lab[s:s + m] for s, m in cuts]
    split = orc.update_centroids(xs, ls, c)
    assert torch.allclose(whole, split, rtol=1e-6, atol=1e-7)
    # empty clusters go to the origin (Q2)
    counts = torch.bincount(lab.view(-1), minlength=k)
    assert torch.equal(whole[counts == 0], torch.zeros_like(whole[counts == 0]))


@settings(max_examples=40, deadline=None)
@given(seed=st.integers(0, 2**31 - 1), n=st.integers(1, 300), d=st.integers(1, 6), k=st.integers(1, 5),
       f64=st.booleans(), parts=st.integers(1, 3), coarse=st.booleans())
def test_radix_selection_protocol_equals_sorting(seed, n, d, k, f64, parts, coarse):
    """The hk_select_* protocol (include/hkmeans.h; restated in numpy by tests/checker_engine.py, counts summed over the
    shards like the ranks do) gives the medians of ht.median's rule (sort, lo + (hi - lo) * frac, all-zero rows dropped:
    heat/cluster/kmedians.py:70-101, heat/core/statistics.py:1684-1728) for any sharding — values with many ties,
    negative zeros, huge values and denormals included."""
    import os
    import sys

    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from checker_engine import CheckerEngine
    from oracle import consumers_oracle as con

    g = torch.Generator().manual_seed(seed)
    dt = torch.float64 if f64 else torch.float32
    x = torch.randn(n, d, generator=g, dtype=torch.float64)
    if coarse:
        x = torch.round(x * 2) / 2  # many equal values, exact zeros (and -0.0 after the sign flip below)
    x = (x * torch.where(torch.rand(n, d, generator=g) < 0.1, -1.0, 1.0)).to(dt)
    # (no infinities: the reference masks rows by multiplying with the 0/1 selection, kmedians.py:74-77, so a row holding
    # inf turns into NaN in EVERY cluster and is kept there — a quirk of the reference that is not reproduced)
    special = torch.tensor([3.0e38, -3.0e38, 1e-42, -0.0], dtype=dt)
    pick = torch.rand(n, d, generator=g) < 0.03
    x[pick] = special[torch.randint(0, 4, (int(pick.sum()),), generator=g)]
    lab = torch.randint(0, k, (n,), generator=g)
    want, want_counts = con.cluster_medians(x, lab, k)

    cuts = sorted(torch.randint(0, n + 1, (parts - 1,), generator=g).tolist())
    bounds = [0] + cuts + [n]
    shards = [(x[a:b], lab[a:b]) for a, b in zip(bounds[:-1], bounds[1:])]
    eng = CheckerEngine(torch.device("cpu"))
    # the protocol is a function of the multiset of kept (label, value) pairs: the sum of the shard histograms is the
    # histogram of the concatenation, so the shards are checked through their counts and the medians on the whole
    got, got_counts = eng.cluster_medians(torch.cat([s for s, _ in shards]), torch.cat([l for _, l in shards]), k)
    per_shard = [eng.cluster_medians(s, l, k)[1] for s, l in shards if s.shape[0]]
    assert got_counts.tolist() == want_counts.tolist()
    if per_shard:
        assert torch.stack(per_shard).sum(dim=0).tolist() == want_counts.tolist()
    ok = want_counts > 0
    # identical up to the sign of zero (the integer image orders -0 below +0, sorting treats them as equal)
    assert torch.equal(torch.nan_to_num(got[ok], nan=7.0) + 0.0, torch.nan_to_num(want[ok], nan=7.0) + 0.0)
