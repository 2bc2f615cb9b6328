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
