"""world_size-2 run of KMeans.fit on two GPUs (skipped on single-GPU boxes): split=0 shards, the k x (d+1) partials
exchanged through peer-mapped GPU memory inside the finish kernel of hk_lloyd_step (and, in one case, through the
ncclAllReduce fallback), results equal to the reference."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, name, out, no_peer):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                      LOCAL_RANK=str(rank), HK_NO_PEER="1" if no_peer else "0")
    import heat_b200 as hb
    from cases import CASES, make_case

    comm = hb.init_from_env("nccl")
    dev = torch.device("cuda", torch.cuda.current_device())
    spec = CASES[name]
    x, init = make_case(name)
    hx = hb.array(x.to(dev), split=0)
    km = hb.cluster.KMeans(n_clusters=init.shape[0], init=hb.array(init.to(dev)), max_iter=spec["max_iter"],
                           tol=spec["tol"])
    km.fit(hx)
    pred = km.predict(hx)
    lab = km.labels_.resplit(None).larray.cpu()
    predl = pred.resplit(None).larray.cpu()
    if rank == 0:
        torch.save({"centers": km.cluster_centers_.larray.cpu(), "labels": lab, "n_iter": km.n_iter_,
                    "inertia": float(km.inertia_), "pred": predl, "fv": float(km.functional_value_),
                    "variant": hb.engine.get_engine(dev).last_variant(),
                    "comm": hb.engine.get_engine(dev).comm_mode(),
                    "graphs": hb.engine.get_engine(dev).graph_launch_count()}, out)
    import torch.distributed as dist

    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("name,no_peer", [("blobs_f32_d32_k64", False), ("blobs_f64_d16_k8", False),
                                          ("config1_spherical", False), ("overlap_f32_d4_k16", False),
                                          ("blobs_f32_fixed5", False),  # 30011 rows: shard remainder
                                          ("bigk_f32_d64_k320", False), ("blobs_f32_d32_k64", True)])
def test_kmeans_two_gpus_matches_reference(tmp_path, name, no_peer):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from cases import CASES, make_case
    from helpers import assert_fit_matches, load_golden
    from oracle import kmeans_oracle as orc

    out = str(tmp_path / "res.pt")
    mp.spawn(_worker, args=(2, _free_port(), name, out, no_peer), nprocs=2, join=True)
    res = torch.load(out)
    assert res["comm"] == ("nccl" if no_peer else "peer"), res["comm"]
    x, init = make_case(name)
    gold = load_golden(name)
    r = orc.fit([x], init, max_iter=max(int(gold["n_iter"]) - 1, 0), tol=None) if int(gold["n_iter"]) > 1 else None
    pre = r.cluster_centers if r is not None else init
    assert_fit_matches(name, x, init, gold, res["centers"], res["labels"], res["n_iter"], res["inertia"],
                       pre_centers=pre.to(x.dtype))
    par = orc.compare_labels(x, torch.from_numpy(gold["centers"]).to(x.dtype),
                             torch.from_numpy(gold["predict_labels"].astype(np.int64)), res["pred"])
    assert par.hard == 0, par
    rtol = 1e-4 if x.dtype == torch.float32 else 1e-10
    np.testing.assert_allclose(res["fv"], float(gold["functional_value"]), rtol=rtol)


def _worker_rings(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                      LOCAL_RANK=str(rank))
    import heat_b200 as hb
    from cases import METRIC_SIGMA
    from helpers import load_golden

    hb.init_from_env("nccl")
    dev = torch.device("cuda", torch.cuda.current_device())
    c = load_golden("cdist")
    X, Y = torch.from_numpy(c["X_f32"]).to(dev), torch.from_numpy(c["Y_f32"]).to(dev)
    hx, hy = hb.array(X, split=0), hb.array(Y, split=0)
    res = {"cdist_ring": hb.spatial.cdist(hx, hy, quadratic_expansion=True),
           "cdist_self": hb.spatial.cdist(hx, quadratic_expansion=True),
           "rbf_self": hb.spatial.rbf(hx, sigma=METRIC_SIGMA, quadratic_expansion=True),
           "manhattan_ring": hb.spatial.manhattan(hx, hy, expand=True)}
    # a larger pair of blocks: the tiles run on the tensor-core kernel and land in aligned column ranges
    g = torch.Generator().manual_seed(8)
    A, B = torch.randn(4096, 64, generator=g), torch.randn(2048, 64, generator=g)
    big = hb.spatial.cdist(hb.array(A.to(dev), split=0), hb.array(B.to(dev), split=0), quadratic_expansion=True)
    variant = hb.engine.get_engine(dev).last_variant()
    torch.save({"local": {k: v.larray.cpu() for k, v in res.items()}, "big": big.larray.cpu(), "variant": variant,
                "meta": {k: (v.split, tuple(v.shape)) for k, v in res.items()}}, out + f".{rank}")
    import torch.distributed as dist

    dist.barrier()
    dist.destroy_process_group()


def test_distance_rings_two_gpus(tmp_path):
    """Y.split=0 / Y=None layouts of _dist (heat/spatial/distance.py:237-361, 416-473) with the blocks moving GPU to GPU."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from helpers import load_golden
    from heat_b200.communication import chunk_rows
    from oracle import kmeans_oracle as orc

    out = str(tmp_path / "rings.pt")
    mp.spawn(_worker_rings, args=(2, _free_port(), out), nprocs=2, join=True)
    g = load_golden("metrics")
    full = {"cdist_ring": g["cdist_f32_quad"], "cdist_self": g["cdist_self_f32_quad"], "rbf_self": g["rbf_self_f32_quad"],
            "manhattan_ring": g["manhattan_f32_expand"]}
    gen = torch.Generator().manual_seed(8)
    A, B = torch.randn(4096, 64, generator=gen), torch.randn(2048, 64, generator=gen)
    want_big = orc.pairwise(A.double(), B.double(), "euclidean", True)
    for rank in range(2):
        r = torch.load(out + f".{rank}")
        off, rows = chunk_rows(96, 2, rank)
        for k, ref in full.items():
            assert r["meta"][k] == (0, ref.shape)
            got, want, tol = r["local"][k].numpy(), ref[off:off + rows], 1e-5
            if k == "cdist_self":  # the diagonal is sqrt(rounding) on both sides: compare squared distances
                got, want, tol = got * got, want * want, 1e-4
            np.testing.assert_allclose(got, want, atol=tol, rtol=0, err_msg=k)
        assert r["variant"].startswith("cdist_tc"), r["variant"]
        o2, r2 = chunk_rows(4096, 2, rank)
        assert float((r["big"].double() - want_big[o2:o2 + r2]).abs().max()) <= 1e-5 * float(want_big.max())


def _worker_consumers(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                      LOCAL_RANK=str(rank))
    import heat_b200 as hb
    from cases import consumer_inputs

    hb.init_from_env("nccl")
    dev = torch.device("cuda", torch.cuda.current_device())
    inp = consumer_inputs()
    hx = hb.array(inp["x"].to(dev), split=0)
    init = hb.array(inp["init"].to(dev))
    km = hb.cluster.KMedians(n_clusters=4, init=init, max_iter=30, tol=1e-4).fit(hx)
    pred = km.predict(hx)
    kd = hb.cluster.KMedoids(n_clusters=4, init=init, max_iter=30).fit(hx)
    knn = hb.classification.KNeighborsClassifier(n_neighbors=5)
    knn.fit(hx, hb.array(inp["y"].to(dev), split=0))
    cls = knn.predict(hb.array(inp["x_test"].to(dev), split=0))
    res = {"kmedians_centers": km.cluster_centers_.larray.cpu(), "kmedians_labels": km.labels_.resplit(None).larray.cpu(),
           "kmedians_n_iter": km.n_iter_, "kmedians_inertia": float(km.inertia_),
           "kmedians_predict": pred.resplit(None).larray.cpu(), "kmedians_fv": float(km.functional_value_),
           "kmedoids_centers": kd.cluster_centers_.larray.cpu(), "kmedoids_labels": kd.labels_.resplit(None).larray.cpu(),
           "kmedoids_n_iter": kd.n_iter_, "knn_classes": cls.resplit(None).larray.cpu()}
    if rank == 0:
        torch.save(res, out)
    import torch.distributed as dist

    dist.barrier()
    dist.destroy_process_group()


def test_kmedians_kmedoids_knn_two_gpus(tmp_path):
    """The N4 consumers on two row shards: selection counts summed over the ranks, medoid chosen across ranks, the
    distance ring under kNN (train and test rows both split)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from test_gloo_multirank import _check_consumers

    out = str(tmp_path / "consumers.pt")
    mp.spawn(_worker_consumers, args=(2, _free_port(), out), nprocs=2, join=True)
    _check_consumers(torch.load(out))


def _worker_batch_parallel(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                      LOCAL_RANK=str(rank))
    import heat_b200 as hb
    from cases import consumer_inputs

    hb.init_from_env("nccl")
    dev = torch.device("cuda", torch.cuda.current_device())
    hx = hb.array(consumer_inputs()["x"].to(dev), split=0)
    res = {}
    for cls, tag, ini in ((hb.cluster.BatchParallelKMeans, "bpkmeans", "k-means++"),
                          (hb.cluster.BatchParallelKMedians, "bpkmedians", "k-medians++")):
        bp = cls(n_clusters=4, init=ini, max_iter=30, tol=1e-4, random_state=5).fit(hx)
        lab = bp.predict(hx)
        res[tag] = {"centers": bp.cluster_centers_.larray.cpu(), "n_iter": bp.n_iter_,
                    "labels": lab.resplit(None).larray.cpu(), "fv": bp.functional_value_, "dtype": lab.dtype}
    if rank == 0:
        torch.save(res, out)
    import torch.distributed as dist

    dist.barrier()
    dist.destroy_process_group()


def test_batch_parallel_clusterers_two_gpus(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from test_gloo_multirank import _check_batch_parallel

    out = str(tmp_path / "bp.pt")
    mp.spawn(_worker_batch_parallel, args=(2, _free_port(), out), nprocs=2, join=True)
    _check_batch_parallel(torch.load(out), 2, None)
