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
