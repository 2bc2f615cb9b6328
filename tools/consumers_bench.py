"""Device-time of the "next" rows on one GPU: rbf / manhattan tiles, one KMedians and one KMedoids
iteration, a kNN prediction.  CUDA events around the public calls, best of 5 after 2 warm-ups; one JSON line."""
import json
import sys

import torch

sys.path.insert(0, ".")
import heat_b200 as hb  # noqa: E402

dev = torch.device("cuda", 0)


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


res = {}
g = torch.Generator(device=dev).manual_seed(1)
# rbf at the config-2 shape (1M x 4096 x 64 fp32, quadratic expansion): same kernel as cdist, Gaussian epilogue
X = hb.array(torch.randn(1_000_000, 64, device=dev, generator=g), split=0)
Y = hb.array(torch.randn(4096, 64, device=dev, generator=g))
eng = hb.engine.get_engine(dev)
out = torch.empty((1_000_000, 4096), device=dev)
res["rbf_1Mx4096x64_ms"] = timed(lambda: eng.pairwise(X.larray, Y.larray, out, "gaussian", True, 8.0))
res["cdist_1Mx4096x64_ms"] = timed(lambda: eng.pairwise(X.larray, Y.larray, out, "euclidean", True))
res["rbf_write_gbs"] = 4.0 * 1e6 * 4096 / res["rbf_1Mx4096x64_ms"] / 1e6
del out
o2 = torch.empty((200_000, 1024), device=dev)
res["manhattan_200kx1024x64_ms"] = timed(lambda: eng.pairwise(X.larray[:200_000], Y.larray[:1024], o2, "manhattan", True))
del o2, X, Y
# KMedians / KMedoids: N = 10M, d = 32, k = 8 (the config-5 shape in fp32)
n, d, k = 10_000_000, 32, 8
cent = 4 * torch.randn(k, d, device=dev, generator=g)
x = cent[torch.randint(0, k, (n,), device=dev, generator=g)] + torch.randn(n, d, device=dev, generator=g)
hx = hb.array(x, split=0)
km = hb.cluster.KMedians(n_clusters=k, init=hb.array(cent + 0.3), max_iter=1, tol=None)
km._initialize_cluster_centers(hx, 2, 1)
lab = km._assign_to_cluster(hx)
res["kmedians_assign_l1_10Mx32x8_ms"] = timed(lambda: km._assign_to_cluster(hx))
res["kmedians_medians_10Mx32x8_ms"] = timed(lambda: km._cluster_medians(hx, lab), reps=3, warm=1)
res["assign_l1_read_gbs"] = 4.0 * n * d / res["kmedians_assign_l1_10Mx32x8_ms"] / 1e6
kd = hb.cluster.KMedoids(n_clusters=k, init=hb.array(cent + 0.3), max_iter=1)
kd._initialize_cluster_centers(hx, 2, 1)
res["kmedoids_update_10Mx32x8_ms"] = timed(lambda: kd._update_centroids(hx, lab), reps=3, warm=1)
# kNN: 20k queries against 100k training rows, d = 32, 5 neighbours
knn = hb.classification.KNeighborsClassifier(n_neighbors=5)
tr = hb.array(x[:100_000].clone(), split=0)
knn.fit(tr, hb.array(torch.randint(0, 10, (100_000,), device=dev, generator=g), split=0))
q = hb.array(x[5_000_000:5_020_000].clone(), split=0)
res["knn_predict_20kx100kx32_ms"] = timed(lambda: knn.predict(q), reps=3, warm=1)
print(json.dumps({k2: round(v, 3) for k2, v in res.items()}))
