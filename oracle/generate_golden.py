"""TEST INFRASTRUCTURE — records outputs of the UNMODIFIED reference for the k-means path.

Run in the build container only (needs /root/reference; the GPU box does not have it):

    python oracle/generate_golden.py            # writes tests/golden/*.npz + MANIFEST.json
    python oracle/generate_golden.py --only NAME   # (re)records one case, keeps the others

The reference is imported from /root/reference under ``oracle/mpi4py_shim`` (mpi4py is not installed in
this image).  Inputs come from seeded generators that the tests re-run (``heat_b200.synthetic`` and the
formulas below); each fixture stores a SHA-256 of its input so a drifting RNG is detected, plus the
reference's outputs: ``cluster_centers_``, ``labels_``, ``n_iter_``, ``inertia_``, and for predict /
cdist the returned arrays.
"""
from __future__ import annotations

import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle", "mpi4py_shim"))
sys.path.insert(1, "/root/reference")
sys.path.insert(2, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def sha(t: torch.Tensor) -> str:
    return hashlib.sha256(t.contiguous().numpy().tobytes()).hexdigest()


# ---- the seeded inputs (re-created verbatim by tests/cases.py) -------------------------------------
sys.path.insert(0, os.path.join(ROOT, "tests"))
from cases import CASES, make_case  # noqa: E402


def run_case(name: str):
    import heat as ht

    spec = CASES[name]
    x, init = make_case(name)
    comm = ht.MPI_WORLD
    hx = ht.array(x, split=spec.get("split", 0))
    hinit = ht.array(init)
    km = ht.cluster.KMeans(n_clusters=init.shape[0], init=hinit, max_iter=spec["max_iter"], tol=spec["tol"])
    km.fit(hx)
    centers = km.cluster_centers_.larray.clone()
    labels = km.labels_.larray.clone()
    out = {
        "centers": centers.numpy(),
        "n_iter": np.int64(km.n_iter_),
        "inertia": km.inertia_.larray.numpy(),
        "x_sha": sha(x),
        "init_sha": sha(init),
    }
    pred = km.predict(hx)
    plab = pred.larray.clone()
    out["functional_value"] = km.functional_value_.larray.numpy()
    # labels travel per rank; gather in rank order through the stand-in's allgather
    if comm.size > 1:
        parts = comm.handle.allgather(labels.numpy())
        labels_all = np.concatenate(parts, axis=0)
        pparts = comm.handle.allgather(plab.numpy())
        plab_all = np.concatenate(pparts, axis=0)
    else:
        labels_all, plab_all = labels.numpy(), plab.numpy()
    k = init.shape[0]
    ldt = np.uint8 if k <= 256 else np.int32
    out["labels"] = labels_all.astype(ldt).reshape(-1)
    out["predict_labels"] = plab_all.astype(ldt).reshape(-1)
    return out


def run_cdist():
    import heat as ht

    out = {}
    g = torch.Generator().manual_seed(7)
    for dt, nm in ((torch.float32, "f32"), (torch.float64, "f64")):
        X = (3 * torch.randn(96, 7, generator=g, dtype=torch.float64)).to(dt)
        Y = (3 * torch.randn(40, 7, generator=g, dtype=torch.float64) + 1).to(dt)
        out[f"X_{nm}"] = X.numpy()
        out[f"Y_{nm}"] = Y.numpy()
        for q in (False, True):
            d = ht.spatial.cdist(ht.array(X, split=0), ht.array(Y), quadratic_expansion=q)
            out[f"D_{nm}_{'quad' if q else 'direct'}"] = d.larray.numpy()
    # the reference's own known-answer test: ones vs zeros in 4-D -> 2.0 (tests/spatial/test_distances.py:14-40)
    d = ht.spatial.cdist(ht.ones((4, 4), split=0), ht.zeros((6, 4)), quadratic_expansion=True)
    out["ones_zeros"] = d.larray.numpy()
    return out


def metric_inputs():
    """Same seeded X (96 x 7) / Y (40 x 7) as run_cdist."""
    out = {}
    g = torch.Generator().manual_seed(7)
    for dt, nm in ((torch.float32, "f32"), (torch.float64, "f64")):
        out[f"X_{nm}"] = (3 * torch.randn(96, 7, generator=g, dtype=torch.float64)).to(dt)
        out[f"Y_{nm}"] = (3 * torch.randn(40, 7, generator=g, dtype=torch.float64) + 1).to(dt)
    return out


METRIC_SIGMA = 2.5


def run_metrics(split_y=None):
    """rbf / manhattan and the Y=None layout through the unmodified reference (heat/spatial/distance.py:159-361).
    Under WORLD_SIZE > 1 (``--metrics-worker``) the distributed layouts run through the reference's rings and every
    rank returns its local block."""
    import heat as ht

    out = {}
    inp = metric_inputs()
    for nm in ("f32", "f64"):
        X, Y = inp[f"X_{nm}"], inp[f"Y_{nm}"]
        hx = ht.array(X, split=0)
        hy = ht.array(Y, split=split_y)
        for q in (False, True):
            tag = "quad" if q else "direct"
            out[f"rbf_{nm}_{tag}"] = ht.spatial.rbf(hx, hy, sigma=METRIC_SIGMA, quadratic_expansion=q).larray.numpy()
            out[f"manhattan_{nm}_{'expand' if q else 'direct'}"] = ht.spatial.manhattan(hx, hy, expand=q).larray.numpy()
            out[f"cdist_{nm}_{tag}"] = ht.spatial.cdist(hx, hy, quadratic_expansion=q).larray.numpy()
            out[f"cdist_self_{nm}_{tag}"] = ht.spatial.cdist(hx, quadratic_expansion=q).larray.numpy()
            out[f"rbf_self_{nm}_{tag}"] = ht.spatial.rbf(hx, sigma=METRIC_SIGMA, quadratic_expansion=q).larray.numpy()
        out[f"manhattan_self_{nm}"] = ht.spatial.manhattan(hx, expand=True).larray.numpy()
    return out


def consumer_inputs():
    """Seeded inputs of the KMedians / KMedoids / kNN goldens (re-created by tests/cases.py:consumer_inputs)."""
    g = torch.Generator().manual_seed(11)
    k, d, n = 4, 5, 1500
    cent = 1.2 * torch.randn(k, d, generator=g, dtype=torch.float64)
    lab = torch.arange(n) % k
    x = (cent[lab] + torch.randn(n, d, generator=g, dtype=torch.float64)).to(torch.float32)
    x[17] = 0.0  # an all-zero row: the reference drops it before taking medians (kmedians.py:76-79)
    init = (cent + 0.5 * torch.randn(k, d, generator=g, dtype=torch.float64)).to(torch.float32)
    xt = (cent[torch.arange(90) % k] + 1.5 * torch.randn(90, d, generator=g, dtype=torch.float64)).to(torch.float32)
    return {"x": x, "init": init, "y": lab.clone(), "x_test": xt}


def run_consumers():
    """KMedians / KMedoids / KNeighborsClassifier through the unmodified reference (heat/cluster/kmedians.py,
    kmedoids.py, heat/classification/kneighborsclassifier.py)."""
    import heat as ht

    inp = consumer_inputs()
    out = {k: v.numpy() for k, v in inp.items()}
    x, init = inp["x"], inp["init"]
    for dt, nm in ((torch.float32, "f32"), (torch.float64, "f64")):
        hx = ht.array(x.to(dt), split=0)
        km = ht.cluster.KMedians(n_clusters=4, init=ht.array(init.to(dt)), max_iter=30, tol=1e-4)
        km.fit(hx)
        out[f"kmedians_{nm}_centers"] = km.cluster_centers_.larray.numpy()
        out[f"kmedians_{nm}_labels"] = km.labels_.larray.numpy()
        out[f"kmedians_{nm}_n_iter"] = np.int64(km.n_iter_)
        out[f"kmedians_{nm}_inertia"] = np.float64(km._inertia.item())
        out[f"kmedians_{nm}_predict"] = km.predict(hx).larray.numpy()
        out[f"kmedians_{nm}_fv"] = np.float64(km.functional_value_.item())
        kd = ht.cluster.KMedoids(n_clusters=4, init=ht.array(init.to(dt)), max_iter=30)
        kd.fit(hx)
        out[f"kmedoids_{nm}_centers"] = kd.cluster_centers_.larray.numpy()
        out[f"kmedoids_{nm}_labels"] = kd.labels_.larray.numpy()
        out[f"kmedoids_{nm}_n_iter"] = np.int64(kd.n_iter_)
        for cls, tag, ini in ((ht.cluster.BatchParallelKMeans, "bpkmeans", "k-means++"),
                              (ht.cluster.BatchParallelKMedians, "bpkmedians", "k-medians++")):
            bp = cls(n_clusters=4, init=ini, max_iter=30, tol=1e-4, random_state=5)
            bp.fit(hx)
            out[f"{tag}_{nm}_centers"] = bp.cluster_centers_.larray.numpy()
            out[f"{tag}_{nm}_n_iter"] = np.int64(bp.n_iter_)
            out[f"{tag}_{nm}_predict"] = bp.predict(hx).larray.numpy()
            out[f"{tag}_{nm}_fv"] = np.float64(bp.functional_value_)
        knn = ht.classification.kneighborsclassifier.KNeighborsClassifier(n_neighbors=5)
        knn.fit(hx, ht.array(inp["y"], split=0))
        out[f"knn_{nm}_classes"] = knn.predict(ht.array(inp["x_test"].to(dt), split=0)).larray.numpy()
    return out


def main():
    os.makedirs(GOLD, exist_ok=True)
    if len(sys.argv) > 1 and sys.argv[1] == "--consumers":
        np.savez_compressed(os.path.join(GOLD, "consumers.npz"), **run_consumers())
        return
    if len(sys.argv) > 1 and sys.argv[1] == "--metrics-worker":
        res = run_metrics(split_y=0)
        np.savez_compressed(os.path.join(GOLD, f"metrics__np{os.environ['WORLD_SIZE']}_rank{os.environ['RANK']}.npz"), **res)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "--metrics":
        np.savez_compressed(os.path.join(GOLD, "metrics.npz"), **run_metrics())
        ws = sys.argv[2] if len(sys.argv) > 2 else "3"
        env = dict(os.environ, WORLD_SIZE=ws, MASTER_ADDR="127.0.0.1", MASTER_PORT="29633", OMP_NUM_THREADS="4")
        procs = [subprocess.Popen([sys.executable, __file__, "--metrics-worker"], env=dict(env, RANK=str(r)))
                 for r in range(int(ws))]
        assert all(p.wait() == 0 for p in procs)
        # the rings must reproduce the rows of the one-process matrices; keep the comparison, not the blocks
        full = np.load(os.path.join(GOLD, "metrics.npz"))
        worst, off = {}, 0
        for r in range(int(ws)):
            fn = os.path.join(GOLD, f"metrics__np{ws}_rank{r}.npz")
            blk = np.load(fn)
            rows = blk["cdist_f32_quad"].shape[0]
            for k in blk.files:
                worst[k] = max(worst.get(k, 0.0), float(np.abs(blk[k] - full[k][off:off + rows]).max()))
            off += rows
            os.remove(fn)
        with open(os.path.join(GOLD, "metrics_rings.json"), "w") as f:
            json.dump({"np": int(ws), "layout": "X.split=0 with Y.split=0 / Y=None through the reference's rings",
                       "max_abs_diff_vs_np1": worst}, f, indent=1, sort_keys=True)
        print("ring layouts through the reference at np=" + ws, "max diff vs np=1:", max(worst.values()), flush=True)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "--rank-worker":
        name = sys.argv[2]
        res = run_case(name)
        if int(os.environ.get("RANK", "0")) == 0:
            np.savez_compressed(os.path.join(GOLD, f"{name}__np{os.environ['WORLD_SIZE']}.npz"), **res)
        return
    import heat as ht

    only = sys.argv[2] if len(sys.argv) > 2 and sys.argv[1] == "--only" else None  # add one case, keep the rest
    manifest = {"reference_version": ht.__version__, "torch": torch.__version__, "cases": {}}
    if only is not None:
        with open(os.path.join(GOLD, "MANIFEST.json")) as f:
            manifest = json.load(f)
    for name, spec in CASES.items():
        if only is not None and name != only:
            continue
        res = run_case(name)
        np.savez_compressed(os.path.join(GOLD, f"{name}.npz"), **res)
        manifest["cases"][name] = {"n_iter": int(res["n_iter"]), "inertia": float(res["inertia"]),
                                   "np": [1]}
        print(name, "n_iter", int(res["n_iter"]), "inertia", float(res["inertia"]), flush=True)
        if spec.get("np2"):
            env = dict(os.environ, WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29631",
                       OMP_NUM_THREADS="4")
            procs = [subprocess.Popen([sys.executable, __file__, "--rank-worker", name],
                                      env=dict(env, RANK=str(r))) for r in range(2)]
            assert all(p.wait() == 0 for p in procs)
            r2 = np.load(os.path.join(GOLD, f"{name}__np2.npz"))
            same = (np.array_equal(r2["centers"], res["centers"]) and np.array_equal(r2["labels"], res["labels"])
                    and int(r2["n_iter"]) == int(res["n_iter"]))
            manifest["cases"][name]["np"].append(2)
            manifest["cases"][name]["np2_bit_identical"] = bool(same)
            print("   np=2 bit-identical to np=1:", same, flush=True)
            if same:
                os.remove(os.path.join(GOLD, f"{name}__np2.npz"))  # Q7: nothing new to store
    if only is None:
        np.savez_compressed(os.path.join(GOLD, "cdist.npz"), **run_cdist())
    with open(os.path.join(GOLD, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
