"""world_size-2 (and 3) gloo runs of the HOST logic on CPU: split=0 sharding, the per-iteration
allreduce, the sticky convergence flag and n_iter bookkeeping — with the oracle standing in for the
device pass (tests/checker_engine.py)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, name, sync_every, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                      LOCAL_RANK=str(rank))
    torch.set_num_threads(2)
    import heat_b200 as hb
    from cases import CASES, make_case
    from checker_engine import CheckerEngine
    from heat_b200 import engine

    comm = hb.init_from_env("gloo")
    assert comm.size == world and comm.rank == rank
    engine.set_engine_factory(lambda dev: CheckerEngine(dev))
    spec = CASES[name]
    x, init = make_case(name)
    hx = hb.array(x, split=0)
    off, rows = hb.communication.chunk_rows(x.shape[0], world, rank)
    assert hx.lshape[0] == rows and hx.shape == tuple(x.shape)
    km = hb.cluster.KMeans(n_clusters=init.shape[0], init=hb.array(init), max_iter=spec["max_iter"],
                           tol=spec["tol"])
    km.sync_every = sync_every
    km.fit(hx)
    pred = km.predict(hx)
    lab = km.labels_.resplit(None).larray
    predl = pred.resplit(None).larray  # collective: every rank takes part
    if rank == 0:
        torch.save({"centers": km.cluster_centers_.larray, "labels": lab, "n_iter": km.n_iter_,
                    "inertia": float(km.inertia_), "pred": predl,
                    "fv": float(km.functional_value_), "split": km.labels_.split,
                    "gshape": km.labels_.shape}, out)
    import torch.distributed as dist

    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,name,sync_every", [(2, "blobs_f32_d8_k6", 1), (2, "blobs_f64_d16_k8", 8),
                                                   (3, "config1_spherical", 8), (2, "overlap_f32_d4_k16", 5)])
def test_kmeans_host_logic_over_gloo(tmp_path, world, name, sync_every):
    from cases import CASES, make_case
    from helpers import assert_fit_matches, load_golden

    out = str(tmp_path / "res.pt")
    mp.spawn(_worker, args=(world, _free_port(), name, sync_every, out), nprocs=world, join=True)
    res = torch.load(out)
    x, init = make_case(name)
    gold = load_golden(name)
    assert res["split"] == 0 and tuple(res["gshape"]) == (x.shape[0], 1)
    assert res["labels"].dtype == torch.int64
    assert_fit_matches(name, x, init, gold, res["centers"], res["labels"], res["n_iter"], res["inertia"])
    assert np.array_equal(res["pred"].view(-1).numpy(), gold["predict_labels"].astype(np.int64))
    np.testing.assert_allclose(res["fv"], float(gold["functional_value"]), rtol=1e-5)


def _worker_random_init(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                      LOCAL_RANK=str(rank))
    torch.set_num_threads(1)
    import heat_b200 as hb

    hb.init_from_env("gloo")
    g = torch.Generator().manual_seed(4)
    xg = torch.randn(20, 3, generator=g)
    # unbalanced shards (7 and 13 rows): built with is_split=0, balanced is unknown
    local = xg[:7] if rank == 0 else xg[7:]
    hx = hb.array(local.clone(), is_split=0)
    assert hx.shape == (20, 3)
    km = hb.cluster.KMeans(n_clusters=5, init="random", random_state=3)
    km._initialize_cluster_centers(hx, 2, 1)
    if rank == 0:
        torch.save({"centers": km.cluster_centers_.larray, "xg": xg}, out)
    import torch.distributed as dist

    dist.barrier()
    dist.destroy_process_group()


def test_random_init_gathers_rows_of_unbalanced_shards(tmp_path):
    """init="random" samples GLOBAL rows; with is_split=0 shards of 7 and 13 rows the row offsets must come from the
    actual local row counts, not from the balanced partition rule (a sampled centroid must be a row of x, never a sum of
    two rows or the zero vector)."""
    out = str(tmp_path / "res.pt")
    mp.spawn(_worker_random_init, args=(2, _free_port(), out), nprocs=2, join=True)
    res = torch.load(out)
    g = torch.Generator()
    g.manual_seed(3)
    idx = torch.randint(0, 19, (5,), generator=g)
    assert torch.equal(res["centers"], res["xg"][idx])


def _worker_kmeanspp(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                      LOCAL_RANK=str(rank))
    torch.set_num_threads(2)
    import heat_b200 as hb
    from checker_engine import CheckerEngine
    from heat_b200 import engine
    from heat_b200.synthetic import blobs_shard, true_centres

    hb.init_from_env("gloo")
    engine.set_engine_factory(lambda dev: CheckerEngine(dev))
    x, _ = blobs_shard(12000, 8, 6, offset=4.0, seed=21)
    hx = hb.array(x, split=0)
    km = hb.cluster.KMeans(n_clusters=6, init="kmeans++", max_iter=50, tol=1e-4, random_state=5)
    assert km.init == "probability_based"  # the reference's alias (kmeans.py:63-64)
    km.fit(hx)
    # the same initialiser under the Manhattan metric (kmedians++ / kmedoids++ aliases)
    kmed = hb.cluster.KMedians(n_clusters=6, init="kmedians++", max_iter=50, tol=1e-4, random_state=5).fit(hx)
    if rank == 0:
        torch.save({"centers": km.cluster_centers_.larray, "true": true_centres(6, 8, 4.0, 21), "n_iter": km.n_iter_,
                    "kmedians_centers": kmed.cluster_centers_.larray, "kmedians_n_iter": kmed.n_iter_}, out)
    import torch.distributed as dist

    dist.barrier()
    dist.destroy_process_group()


def test_kmeanspp_init_host_logic_over_gloo(tmp_path):
    """init="kmeans++" (k-means||, _kcluster.py:146-245): sampling rounds, candidate gathering across ranks, weights and
    reclustering run identically on both ranks; on well-separated blobs the fit then finds every true centre."""
    out = str(tmp_path / "res.pt")
    mp.spawn(_worker_kmeanspp, args=(2, _free_port(), out), nprocs=2, join=True)
    res = torch.load(out)
    dist = torch.cdist(res["true"].double(), res["centers"].double())
    assert float(dist.min(dim=1).values.max()) < 0.2, dist.min(dim=1).values
    assert res["n_iter"] <= 20
    dist = torch.cdist(res["true"].double(), res["kmedians_centers"].double())
    assert float(dist.min(dim=1).values.max()) < 0.3 and res["kmedians_n_iter"] <= 30, dist.min(dim=1).values


def _worker_rings(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                      LOCAL_RANK=str(rank))
    torch.set_num_threads(1)
    import heat_b200 as hb
    from cases import METRIC_SIGMA
    from checker_engine import CheckerEngine
    from heat_b200 import engine
    from helpers import load_golden

    comm = hb.init_from_env("gloo")
    engine.set_engine_factory(lambda dev: CheckerEngine(dev))
    c = load_golden("cdist")
    X, Y = torch.from_numpy(c["X_f32"]), torch.from_numpy(c["Y_f32"])
    res = {}
    hx, hy = hb.array(X, split=0), hb.array(Y, split=0)
    res["cdist_ring"] = hb.spatial.cdist(hx, hy, quadratic_expansion=True)          # distance.py:416-473
    res["cdist_self"] = hb.spatial.cdist(hx, quadratic_expansion=True)              # distance.py:237-361
    res["rbf_self"] = hb.spatial.rbf(hx, sigma=METRIC_SIGMA, quadratic_expansion=True)
    res["manhattan_ring"] = hb.spatial.manhattan(hx, hy, expand=True)
    res["split1"] = hb.spatial.cdist(hb.array(X), hy, quadratic_expansion=True)     # X replicated, Y split: split=1
    # unbalanced blocks, one of them empty: the counts come from the actual local shapes
    cut = [0, 50, 50, 96][: world + 1] if world == 3 else [0, 70, 96]
    ux = hb.array(X[cut[rank]:cut[rank + 1]].clone(), is_split=0)
    ycut = [0, 0, 25, 40][: world + 1] if world == 3 else [0, 11, 40]
    uy = hb.array(Y[ycut[rank]:ycut[rank + 1]].clone(), is_split=0)
    res["unbalanced"] = hb.spatial.cdist(ux, uy, quadratic_expansion=True)
    meta = {k: (v.split, tuple(v.shape), tuple(v.lshape)) for k, v in res.items()}
    torch.save({"local": {k: v.larray for k, v in res.items()}, "meta": meta, "cut": cut}, out + f".{rank}")
    import torch.distributed as dist

    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_distance_rings_over_gloo(tmp_path, world):
    """Y.split=0 and Y=None layouts of _dist (heat/spatial/distance.py:237-361, 416-473) and the split=1 result of a
    replicated X: every rank's block equals the rows (or columns) of the reference's one-process matrix."""
    from cases import METRIC_SIGMA  # noqa: F401
    from helpers import load_golden

    out = str(tmp_path / "rings.pt")
    mp.spawn(_worker_rings, args=(world, _free_port(), out), nprocs=world, join=True)
    g = load_golden("metrics")
    full = {"cdist_ring": g["cdist_f32_quad"], "cdist_self": g["cdist_self_f32_quad"], "rbf_self": g["rbf_self_f32_quad"],
            "manhattan_ring": g["manhattan_f32_expand"], "unbalanced": g["cdist_f32_quad"]}
    from heat_b200.communication import chunk_rows

    for rank in range(world):
        r = torch.load(out + f".{rank}")
        off, rows = chunk_rows(96, world, rank)
        for k, ref in full.items():
            split, gshape, lshape = r["meta"][k]
            lo, hi = (r["cut"][rank], r["cut"][rank + 1]) if k == "unbalanced" else (off, off + rows)
            assert split == 0 and gshape == ref.shape and lshape == (hi - lo, ref.shape[1]), (k, r["meta"][k])
            np.testing.assert_allclose(r["local"][k].numpy(), ref[lo:hi], atol=1e-5, rtol=0, err_msg=k)
        split, gshape, lshape = r["meta"]["split1"]
        yo, yr = chunk_rows(40, world, rank)
        assert split == 1 and gshape == (96, 40) and lshape == (96, yr)
        np.testing.assert_allclose(r["local"]["split1"].numpy(), g["cdist_f32_quad"][:, yo:yo + yr], atol=1e-5, rtol=0)


def _worker_consumers(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                      LOCAL_RANK=str(rank))
    torch.set_num_threads(1)
    import heat_b200 as hb
    from cases import consumer_inputs
    from checker_engine import CheckerEngine
    from heat_b200 import engine

    hb.init_from_env("gloo")
    engine.set_engine_factory(lambda dev: CheckerEngine(dev))
    inp = consumer_inputs()
    hx = hb.array(inp["x"], split=0)
    km = hb.cluster.KMedians(n_clusters=4, init=hb.array(inp["init"]), max_iter=30, tol=1e-4).fit(hx)
    pred = km.predict(hx)
    kd = hb.cluster.KMedoids(n_clusters=4, init=hb.array(inp["init"]), max_iter=30).fit(hx)
    knn = hb.classification.KNeighborsClassifier(n_neighbors=5)
    knn.fit(hx, hb.array(inp["y"], split=0))
    cls = knn.predict(hb.array(inp["x_test"], split=0))
    res = {"kmedians_centers": km.cluster_centers_.larray, "kmedians_labels": km.labels_.resplit(None).larray,
           "kmedians_n_iter": km.n_iter_, "kmedians_inertia": float(km.inertia_), "kmedians_predict": pred.resplit(None).larray,
           "kmedians_fv": float(km.functional_value_), "kmedoids_centers": kd.cluster_centers_.larray,
           "kmedoids_labels": kd.labels_.resplit(None).larray, "kmedoids_n_iter": kd.n_iter_,
           "knn_classes": cls.resplit(None).larray, "knn_split": cls.split}
    if rank == 0:
        torch.save(res, out)
    import torch.distributed as dist

    dist.barrier()
    dist.destroy_process_group()


def _check_consumers(res, nm="f32"):
    from helpers import load_golden

    g = load_golden("consumers")
    assert res["kmedians_n_iter"] == int(g[f"kmedians_{nm}_n_iter"])
    assert np.array_equal(res["kmedians_labels"].cpu().numpy(), g[f"kmedians_{nm}_labels"])
    np.testing.assert_allclose(res["kmedians_centers"].cpu().numpy(), g[f"kmedians_{nm}_centers"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(res["kmedians_inertia"], float(g[f"kmedians_{nm}_inertia"]), atol=1e-10)
    assert np.array_equal(res["kmedians_predict"].cpu().numpy(), g[f"kmedians_{nm}_predict"])
    np.testing.assert_allclose(res["kmedians_fv"], float(g[f"kmedians_{nm}_fv"]), rtol=1e-5)
    assert res["kmedoids_n_iter"] == int(g[f"kmedoids_{nm}_n_iter"])
    assert np.array_equal(res["kmedoids_labels"].cpu().numpy(), g[f"kmedoids_{nm}_labels"])
    assert np.array_equal(res["kmedoids_centers"].cpu().numpy(), g[f"kmedoids_{nm}_centers"])
    assert np.array_equal(res["knn_classes"].cpu().numpy(), g[f"knn_{nm}_classes"])


@pytest.mark.parametrize("world", [2, 3])
def test_kmedians_kmedoids_knn_host_logic_over_gloo(tmp_path, world):
    """split=0 shards: the radix-selection protocol (counts summed over the ranks), the cross-rank medoid choice, the
    distance ring under kNN — results equal to the reference's one-process goldens (tests/golden/consumers.npz)."""
    out = str(tmp_path / "consumers.pt")
    mp.spawn(_worker_consumers, args=(world, _free_port(), out), nprocs=world, join=True)
    res = torch.load(out)
    assert res["knn_split"] == 0
    _check_consumers(res)


def _worker_batch_parallel(rank, world, port, out, merge):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                      LOCAL_RANK=str(rank))
    torch.set_num_threads(1)
    import heat_b200 as hb
    from cases import consumer_inputs
    from checker_engine import CheckerEngine
    from heat_b200 import engine

    hb.init_from_env("gloo")
    engine.set_engine_factory(lambda dev: CheckerEngine(dev))
    hx = hb.array(consumer_inputs()["x"], split=0)
    res = {}
    for cls, tag, ini in ((hb.cluster.BatchParallelKMeans, "bpkmeans", "k-means++"),
                          (hb.cluster.BatchParallelKMedians, "bpkmedians", "k-medians++")):
        bp = cls(n_clusters=4, init=ini, max_iter=30, tol=1e-4, random_state=5, n_procs_to_merge=merge).fit(hx)
        lab = bp.predict(hx)
        res[tag] = {"centers": bp.cluster_centers_.larray, "n_iter": bp.n_iter_, "labels": lab.resplit(None).larray,
                    "fv": bp.functional_value_, "dtype": lab.dtype}
    # KMeans(init="batchparallel") takes the batch-parallel centres (max_iter=100) as its start (_kcluster.py:249-275)
    km = hb.cluster.KMeans(n_clusters=4, init="batchparallel", random_state=5)
    km._initialize_cluster_centers(hx, 2, 1)
    ref = hb.cluster.BatchParallelKMeans(n_clusters=4, init="k-means++", max_iter=100, random_state=5).fit(hx)
    assert torch.equal(km.cluster_centers_.larray, ref.cluster_centers_.larray)
    if rank == 0:
        torch.save(res, out)
    import torch.distributed as dist

    dist.barrier()
    dist.destroy_process_group()


def _check_batch_parallel(res, world, merge, nm="f32"):
    """world = 1: against the unmodified reference's golden; more ranks: against the restated algorithm run over the same
    shards in one process (the reference's sub-communicator merge cannot run under the mpi4py stand-in)."""
    from cases import consumer_inputs
    from heat_b200.communication import chunk_rows
    from helpers import load_golden
    from oracle import consumers_oracle as con

    dt = torch.float32 if nm == "f32" else torch.float64
    x = consumer_inputs()["x"].to(dt)
    g = load_golden("consumers")
    for p, tag in ((2, "bpkmeans"), (1, "bpkmedians")):
        r = res[tag]
        if world == 1:
            want_c, want_it = torch.from_numpy(g[f"{tag}_{nm}_centers"]), int(g[f"{tag}_{nm}_n_iter"])
            want_lab, want_fv = torch.from_numpy(g[f"{tag}_{nm}_predict"]), float(g[f"{tag}_{nm}_fv"])
        else:
            shards = [x[o:o + c] for o, c in (chunk_rows(x.shape[0], world, rk) for rk in range(world))]
            want_c, want_it = con.batch_parallel_fit(shards, p, 4, 30, 1e-4, 5, merge)
            want_lab, want_fv = con.batch_parallel_predict(x, want_c, p)
        assert r["n_iter"] == want_it, (tag, r["n_iter"], want_it)
        np.testing.assert_allclose(r["centers"].cpu().numpy(), want_c.numpy(), rtol=2e-5, atol=2e-5, err_msg=tag)
        assert r["dtype"] == torch.int32
        assert int((r["labels"].cpu() != want_lab).sum()) <= 2, tag  # rows at the rounding of a boundary
        np.testing.assert_allclose(r["fv"], want_fv, rtol=1e-4)


@pytest.mark.parametrize("world,merge", [(2, None), (3, 2)])
def test_batch_parallel_clusterers_over_gloo(tmp_path, world, merge):
    """BatchParallelKMeans / KMedians (heat/cluster/batchparallelclustering.py:171-331): per-rank clustering, hierarchical
    merge of the centres (3 ranks merged 2 at a time: two levels), centres of rank 0 everywhere, int32 labels."""
    out = str(tmp_path / "bp.pt")
    mp.spawn(_worker_batch_parallel, args=(world, _free_port(), out, merge), nprocs=world, join=True)
    _check_batch_parallel(torch.load(out), world, merge)
