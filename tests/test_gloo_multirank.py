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
    if rank == 0:
        torch.save({"centers": km.cluster_centers_.larray, "true": true_centres(6, 8, 4.0, 21), "n_iter": km.n_iter_}, out)
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
