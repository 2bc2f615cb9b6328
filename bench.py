#!/usr/bin/env python
"""Benchmark of the k-means Lloyd hot path (BASELINE.json metric: Lloyd iter/s at 100M x 32, k=64).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload config3|config5|config2|config4]
                    [--data blobs|overlap|randn|uncentred]

* a "step" is one Lloyd iteration (fused assign + accumulate pass, allreduce of the k x (d+1) partials,
  finalize) over the whole 100M-row matrix, sharded split=0 over N ranks (strong scaling).
* ``value``: iterations/s with X resident in HBM, CUDA events on the launch stream, max over ranks.
* ``e2e``: the same iteration driven from HOST buffers through the public API / C ABI: per step the shard
  is copied from pinned host memory to the device, one step runs, centroids + shift come back.
* ``roofline``: algorithmic bytes (N*d*4 per iteration) / mean duration of the pass kernel (event pairs
  recorded inside the library around that kernel) against MEASURED_PEAKS.json's HBM copy bandwidth.
* ``cpu_baseline`` / ``--impl reference``: the UNMODIFIED reference (``heat.cluster.KMeans`` / ``ht.spatial.cdist`` from
  ``baseline/_ref`` under ``oracle/mpi4py_shim``; torch CPU, all host threads) on a bounded row sample, extrapolated
  linearly in N (stated in ``sample``); the oracle port only if the reference cannot be imported.
* ``parity``: after the timed loop every rank checks that its centroids are bit-equal to rank 0's, and a fixed
  200 003-row problem (a row count no rank count divides) goes through the same sharded path for 3 steps and is compared with
  the oracle on rank 0; a mismatch fails the run.
* ``filter``: what fraction of rows the TF32 filter of the tensor-core pass could not decide (exact re-evaluation).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

WORKLOADS = {
    # BASELINE.json configs[2] — the configuration the metric is quoted on
    "config3": dict(kind="kmeans", n=100_000_000, d=32, k=64, dtype="f32",
                    desc="KMeans N=100M d=32 k=64 fp32 split=0 (BASELINE configs[2])"),
    "config5": dict(kind="kmeans", n=50_000_000, d=16, k=8, dtype="f64",
                    desc="KMeans N=50M d=16 k=8 fp64 split=0 (BASELINE configs[4])"),
    "config4": dict(kind="kmeans", n=20_000_000, d=128, k=1024, dtype="f32",
                    desc="KMeans N=20M d=128 k=1024 fp32 split=0 (BASELINE configs[3])"),
    "config2": dict(kind="cdist", n=1_000_000, d=64, k=4096, dtype="f32",
                    desc="cdist X 1Mx64 vs Y 4096x64 fp32 quadratic_expansion (BASELINE configs[1])"),
}
DT = {"f32": torch.float32, "f64": torch.float64}


def measured_traffic(workload: str, variant: str):
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture (or None)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(p):
        return None
    try:
        t = json.load(open(p))
    except Exception:
        return None
    ent = t.get(workload)
    if not ent:
        return None
    fam = variant.split("<")[0]
    return ent.get(fam)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return float(j["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)", j
    return 6650.0, "fallback (B200_PROFILING.md)", {}


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 200 ms while the timed region runs."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            for nm, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


SAMPLE_ROWS = {"config3": 250_000, "config5": 1_000_000, "config4": 8_000, "config2": 20_000}


def _import_reference():
    """The unmodified reference from baseline/_ref under the mpi4py stand-in, or None."""
    shim = os.path.join(ROOT, "oracle", "mpi4py_shim")
    ref = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref, "heat")):
        return None
    for q in (ref, shim):
        if q not in sys.path:
            sys.path.insert(0, q)
    try:
        import heat as ht

        return ht
    except Exception:
        return None


def timed_cpu_reference(n_sample: int, d: int, k: int, dtype, steps: int, warmup: int, kind: str, data: str = "blobs"):
    """One step of the reference's own CPU implementation on a bounded sample -> (sec/step, threads, kind)."""
    from heat_b200.synthetic import dataset_init, dataset_shard

    # all host cores, also under torchrun (which exports OMP_NUM_THREADS=1 to every rank; only rank 0 runs this)
    try:
        torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    except (AttributeError, RuntimeError):
        torch.set_num_threads(max(1, os.cpu_count() or 1))
    x, _ = dataset_shard(data, n_sample, d, k if kind == "kmeans" else 16, dtype=dtype)
    ht = _import_reference()
    if ht is not None:
        impl = "reference"
        hx = ht.array(x, split=0)
        if kind == "cdist":
            hy = ht.array(torch.randn(k, d, dtype=dtype))
            fn = lambda: ht.spatial.cdist(hx, hy, quadratic_expansion=True)
        else:
            hc = ht.array(dataset_init(data, k, d, dtype=dtype))
            # one Lloyd iteration = fit(max_iter=1, tol=None) with supplied centroids (kmeans.py:131-144)
            fn = lambda: ht.cluster.KMeans(n_clusters=k, init=hc, max_iter=1, tol=None).fit(hx)
    else:
        from oracle import kmeans_oracle as orc

        impl = "port"
        if kind == "cdist":
            y = torch.randn(k, d, dtype=dtype)
            fn = lambda: orc.cdist(x, y, quadratic_expansion=True)
        else:
            c = dataset_init(data, k, d, dtype=dtype)

            def fn():
                lab = orc.assign_to_cluster(x, c)
                new = orc.update_centroids([x], [lab], c)
                return ((c - new) ** 2).sum()

    for _ in range(warmup):
        fn()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    return (time.perf_counter() - t0) / max(steps, 1), torch.get_num_threads(), impl


def reference_on_gpu(n_rows: int, d: int, k: int, dtype, dev, data: str):
    """Second bar: the unmodified reference on one B200 through torch CUDA (np=1), or None."""
    ht = _import_reference()
    if ht is None:
        return None
    from heat_b200.synthetic import dataset_init, dataset_shard

    try:
        x, _ = dataset_shard(data, n_rows, d, k, device=dev, dtype=dtype)
        hx = ht.array(x, split=0, device="gpu")
        hc = ht.array(dataset_init(data, k, d, dtype=dtype).to(dev), device="gpu")
        fn = lambda: ht.cluster.KMeans(n_clusters=k, init=hc, max_iter=1, tol=None).fit(hx)
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        sec = (time.perf_counter() - t0) / reps
        del hx, x
        torch.cuda.empty_cache()
        return {"rows": n_rows, "sec_per_iter": sec,
                "what": "unmodified heat.cluster.KMeans (baseline/_ref), torch CUDA on this B200, np=1, fit(max_iter=1)"}
    except Exception as ex:  # out of memory etc.: reported, not hidden
        torch.cuda.empty_cache()
        return {"rows": n_rows, "error": str(ex)[:160]}


def run_reference(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # rank 0 alone runs the CPU arm
    n, d, k, dtype = wl["n"], wl["d"], wl["k"], DT[wl["dtype"]]
    # bounded sample: about 1-2 s of CPU work per step (k full passes in fp64)
    n_sample = SAMPLE_ROWS[args.workload]
    sec, thr, impl = timed_cpu_reference(n_sample, d, k, dtype, args.steps, args.warmup, wl["kind"], args.data)
    scale = n / n_sample
    value = 1.0 / (sec * scale)
    what = ("unmodified reference (heat 1.9.0-dev from baseline/_ref under oracle/mpi4py_shim, torch CPU"
            if impl == "reference" else "oracle port of the reference (torch CPU")
    sample = (f"{what}, {thr} threads) on {n_sample} of {n} rows, {sec * 1e3:.1f} ms/step, "
              f"extrapolated linearly in N (x{scale:.0f})")
    line = {
        "impl": "reference", "metric": metric_name(wl), "value": value, "unit": unit_name(wl),
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * scale * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": wl["dtype"],
        "data": "synthetic", "config": {"workload": wl["desc"], "n_sample": n_sample, "dataset": args.data},
        "cpu_baseline": {"value": value, "unit": unit_name(wl), "cores": thr, "kind": impl, "sample": sample},
        "e2e": {"value": value, "unit": unit_name(wl), "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "host_cores": os.cpu_count(),
    }
    print(json.dumps(line), flush=True)


def metric_name(wl):
    return "KMeans Lloyd iter/s" if wl["kind"] == "kmeans" else "cdist calls/s"


def unit_name(wl):
    return "iter/s" if wl["kind"] == "kmeans" else "calls/s"


def parity_check(eng, world, rank, dev, wl, args, c_timed):
    """(i) the centroids of the timed run are bit-equal on every rank; (ii) a fixed 200 003-row problem (not divisible
    by any rank count used) goes through the same sharded path for 3 steps and matches the oracle on rank 0 within the
    north-star tolerance.  Raises on a mismatch (the run fails)."""
    import torch.distributed as dist

    from heat_b200.synthetic import dataset_init, dataset_shard

    k, d, dtype = wl["k"], wl["d"], DT[wl["dtype"]]
    res = {"ranks_bit_equal": True}
    if world > 1:
        parts = [torch.empty_like(c_timed) for _ in range(world)]
        dist.all_gather(parts, c_timed.contiguous())
        res["ranks_bit_equal"] = all(torch.equal(p.view(torch.uint8), parts[0].view(torch.uint8)) for p in parts)
    n_par = 200_003
    # generated on the host: the CPU and CUDA generators give different streams, and rank 0 rebuilds the global matrix
    x, _ = dataset_shard("blobs", n_par, d, k, rank, world, device="cpu", dtype=dtype, seed=5)
    x = x.to(dev)
    c = dataset_init("blobs", k, d, dtype=dtype, seed=5).to(dev)
    cp, sh = torch.empty_like(c), torch.zeros((), dtype=dtype, device=dev)
    st = torch.zeros(4, dtype=torch.int32, device=dev)
    ws = eng.row_workspace(x.shape[0])
    for _ in range(3):
        eng.lloyd_step(x, c, cp, False, 0.0, sh, st, world > 1, path=args.path, row_ws=ws)
    torch.cuda.synchronize()
    if world > 1:
        parts = [torch.empty_like(c) for _ in range(world)]
        dist.all_gather(parts, c)
        res["ranks_bit_equal"] = res["ranks_bit_equal"] and all(
            torch.equal(p.view(torch.uint8), parts[0].view(torch.uint8)) for p in parts)
    res.update({"rows": n_par, "steps": 3, "rows_per_rank": int(x.shape[0]), "comm": eng.comm_mode()})
    if rank == 0:
        from oracle import kmeans_oracle as orc  # the checker, never the thing measured

        xf, _ = dataset_shard("blobs", n_par, d, k, device="cpu", dtype=dtype, seed=5)
        cc = dataset_init("blobs", k, d, dtype=dtype, seed=5)
        for _ in range(3):
            lab = orc.assign_to_cluster(xf, cc)
            cc = orc.update_centroids_fast([xf], [lab], cc)
        err = orc.centers_rel_err(cc, c.cpu())
        tol = 1e-5 if dtype == torch.float32 else 1e-12
        res.update({"centroid_rel_err_vs_oracle": err, "tol": tol, "ok": bool(err <= tol and res["ranks_bit_equal"])})
    else:
        res["ok"] = bool(res["ranks_bit_equal"])
    ok = torch.tensor([1 if res["ok"] else 0], device=dev)
    if world > 1:
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if int(ok.item()) != 1:
        raise SystemExit(f"bench.py parity check FAILED on rank {rank}: {res}")
    return res


def run_ours(args, wl):
    import torch.distributed as dist

    import heat_b200 as hb
    from heat_b200.synthetic import dataset_init, dataset_shard

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local % torch.cuda.device_count())
    dev = torch.device("cuda", torch.cuda.current_device())
    comm = hb.init_from_env("nccl") if world > 1 else hb.get_comm()
    n, d, k, dtype = wl["n"], wl["d"], wl["k"], DT[wl["dtype"]]
    esz = 4 if dtype == torch.float32 else 8
    eng = hb.engine.get_engine(dev)
    if world > 1:
        eng.init_comm(comm)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v: float) -> float:
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(v: float) -> float:
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    sampler = ClockSampler(torch.cuda.current_device())

    row_ws = None
    if wl["kind"] == "kmeans":
        x, off = dataset_shard(args.data, n, d, k, rank, world, device=dev, dtype=dtype)
        c0 = dataset_init(args.data, k, d, dtype=dtype).to(dev)
        c = c0.clone()
        c_prev = torch.empty_like(c)
        shift2 = torch.zeros((), dtype=dtype, device=dev)
        state = torch.zeros(4, dtype=torch.int32, device=dev)
        n_loc = x.shape[0]
        # per-matrix workspace (|x| bounds of the tensor-core pass), exactly as KMeans.fit allocates it
        row_ws = None if args.no_row_ws else eng.row_workspace(n_loc)

        def step():
            eng.lloyd_step(x, c, c_prev, False, 0.0, shift2, state, world > 1, path=args.path, row_ws=row_ws)

        alg_bytes_rank = n_loc * d * esz
    else:
        off, n_loc = hb.communication.chunk_rows(n, world, rank)
        g = torch.Generator(device=dev).manual_seed(1 + rank)
        x = torch.randn(n_loc, d, generator=g, device=dev, dtype=dtype)
        y = torch.randn(k, d, generator=torch.Generator(device=dev).manual_seed(7), device=dev, dtype=dtype)
        out = torch.empty((n_loc, k), dtype=dtype, device=dev)

        def step():
            eng.cdist(x, y, out, quadratic_expansion=True)

        alg_bytes_rank = esz * (n_loc * k + n_loc * d + k * d)

    # ---- device-resident timing -----------------------------------------------------------------------
    # k-means: the timed region is ONE hk_lloyd_run call of K iterations - what KMeans.fit issues between two reads of the
    # convergence flag (first iteration eager, the rest replayed from CUDA graphs; 2 kernels per iteration).  The dominant
    # kernel's own duration comes from a second region of K eager iterations with an event pair around every launch of it
    # (events cannot be recorded inside a replayed graph, and they cost ~20 us per iteration, which is why they are kept
    # out of the headline region).  Both regions start cold (row workspace zeroed) as a fit does.
    kmeans = wl["kind"] == "kmeans"

    def run_k(nsteps):
        if kmeans:
            eng.lloyd_run(x, c, c_prev, False, 0.0, shift2, state, world > 1, nsteps, path=args.path, row_ws=row_ws)
        else:
            for _ in range(nsteps):
                step()

    # the clock sampler (an nvidia-smi process per rank) starts BEFORE the warm-up: its start-up holds driver locks for
    # tens of ms, which at 8 GPUs is several times the whole timed region; it then samples every 100 ms through the
    # warm-up, the timed region and the profiled region (all under load)
    sampler.start()
    run_k(args.warmup)
    time.sleep(0.3)
    barrier()
    if row_ws is not None:
        row_ws.zero_()
    eng.stats()  # clear the cold-path counters
    barrier()
    l0 = eng.launch_count()
    gl0 = eng.graph_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run_k(args.steps)
    e1.record()
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    launches = sum_over_ranks(eng.launch_count() - l0)
    graph_launches = eng.graph_launch_count() - gl0
    ms_step = ms_total / args.steps
    value = 1e3 / ms_step
    filt = None
    parity = None
    if kmeans:
        iters_done = int(state.cpu()[1])
        assert iters_done == args.warmup + args.steps, (iters_done, args.warmup + args.steps)
        st = eng.stats()
        if st["passes"]:
            filt = {"undecided_frac": st["undecided_frac"], "undecided_rows_per_pass": st["undecided_rows"] / st["passes"],
                    "exact_pairs_per_undecided_row": (st["exact_pairs"] / st["undecided_rows"]) if st["undecided_rows"] else 0.0,
                    "all_centroid_rows": st["all_centroid_rows"], "passes": st["passes"],
                    "what": "rows of the timed passes the TF32 filter could not decide (re-evaluated with the exact formula)"}
    # second region: K eager iterations with per-kernel events.  A pause first: at ~1 kW this kernel runs into the board's
    # power cap within ~0.1 s of sustained load (sw_power_cap, clocks drop ~10 %); both regions are K-step bursts from an
    # idle GPU, which is also how MEASURED_PEAKS.json's copy bandwidth (best of 10 short copies) was taken.
    if row_ws is not None:
        row_ws.zero_()
    barrier()
    time.sleep(1.0)
    barrier()
    eng.profile(True)
    eng.profile_read()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    for _ in range(args.steps):
        step()
    p1.record()
    barrier()
    prof_ms_step = max_over_ranks(p0.elapsed_time(p1)) / args.steps
    clocks = sampler.stop()
    kern_ms, kern_n = eng.profile_read()
    eng.profile(False)
    variant = eng.last_variant()
    kern_ms_avg = max_over_ranks(kern_ms / max(kern_n, 1))
    graph = {"graph_launches": graph_launches, "eager_profiled_ms_per_step": prof_ms_step,
             "what": "timed region = one hk_lloyd_run call (first iteration eager, the rest replayed from CUDA graphs); "
                     "eager_profiled_ms_per_step = the second region (per-kernel events, plain launches)"} if kmeans else None
    if kmeans:
        parity = parity_check(eng, world, rank, dev, wl, args, c)

    # ---- end to end from host buffers -------------------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        try:
            xh = torch.empty((n_loc, d), dtype=dtype, pin_memory=True)
            xh.copy_(x)
            torch.cuda.synchronize()
            if wl["kind"] == "kmeans":
                ch = torch.empty((k, d), dtype=dtype, pin_memory=True)
                sh = torch.empty((), dtype=dtype, pin_memory=True)
                c.copy_(c0)

                cbox = [c0.clone()]

                def e2e_step():
                    # the call a user makes: host rows -> device DNDarray -> KMeans.fit (one Lloyd iteration, labels,
                    # inertia) -> centroids back on the host
                    x.copy_(xh, non_blocking=True)
                    xd = hb.dndarray.DNDarray(x, (n, d), x.dtype, 0, dev, comm, True)
                    km = hb.cluster.KMeans(n_clusters=k, init=hb.array(cbox[0]), max_iter=1, tol=None)
                    km.kernel_path = args.path
                    km.fit(xd)
                    cbox[0] = km.cluster_centers_.larray
                    ch.copy_(cbox[0], non_blocking=True)
                    sh.copy_(km.inertia_.larray, non_blocking=True)

                d2h = k * d * esz + esz
            else:
                oh = torch.empty((n_loc, k), dtype=dtype, pin_memory=True)

                def e2e_step():
                    x.copy_(xh, non_blocking=True)
                    eng.cdist(x, y, out, quadratic_expansion=True)
                    oh.copy_(out, non_blocking=True)

                d2h = n_loc * k * esz
            e2e_steps = args.steps
            for _ in range(min(args.warmup, 3)):
                e2e_step()
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(e2e_steps):
                e2e_step()
            b.record()
            barrier()
            e2e_ms = max_over_ranks(a.elapsed_time(b)) / e2e_steps
            e2e = {"value": 1e3 / e2e_ms, "unit": unit_name(wl), "ms_per_step": e2e_ms,
                   "h2d_bytes_per_step": int(sum_over_ranks(n_loc * d * esz)),
                   "d2h_bytes_per_step": int(sum_over_ranks(d2h)), "steps": e2e_steps,
                   "api": ("heat_b200.cluster.KMeans(init=centroids, max_iter=1, tol=None).fit(DNDarray) -> hk_lloyd_run + "
                           "hk_assign (C ABI); pinned host X copied to the device every step, centroids + inertia read back"
                           if wl["kind"] == "kmeans" else
                           "heat_b200 engine.cdist -> hk_cdist (C ABI); pinned host X copied every step, distances read back")}
            del xh
        except RuntimeError as ex:  # pinned allocation failure is reported, not hidden
            e2e = {"value": None, "unit": unit_name(wl), "error": str(ex)[:200], "h2d_bytes_per_step": 0,
                   "d2h_bytes_per_step": 0}

    if rank != 0:
        return
    peak, peak_src, pk = peaks()
    achieved = alg_bytes_rank / (kern_ms_avg * 1e-3) / 1e9 if kern_ms_avg > 0 else 0.0
    roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": measured_traffic(args.workload, variant) if world == 1 and not args.rows else None,
            "peak_source": peak_src, "kernel": variant, "kernel_ms_avg": kern_ms_avg,
            "algorithmic_bytes_per_launch": alg_bytes_rank,
            "kernel_share_of_step": kern_ms_avg / prof_ms_step if prof_ms_step > 0 else None,
            "measured_in": "second K-step region of this run: eager launches, one CUDA-event pair around every launch of this kernel"}
    if variant.startswith("bigk<"):
        # large-k path: tensor-core bound.  The dominant kernel is the tcgen05 3xTF32 distance
        # kernel, launched once per row chunk; algorithmic FLOPs per launch = 2 * rows * k * d (it executes 3x that).
        launches_per_step = kern_n / max(args.steps, 1)
        alg_flops_launch = 2.0 * n_loc * k * d / max(launches_per_step, 1)
        tf32_peak = float(pk.get("bf16_tflops_sustained", 1344.2)) / 2.0
        ach = alg_flops_launch / (kern_ms_avg * 1e-3) / 1e12 if kern_ms_avg > 0 else 0.0
        roof = {"bound": "tensor", "achieved": ach, "peak": tf32_peak, "unit": "TFLOP/s", "frac": ach / tf32_peak,
                "traffic": None,
                "peak_source": "derived: MEASURED_PEAKS.json bf16_tflops_sustained / 2 (dense TF32; not measured directly)",
                "kernel": variant + " -> cdist_tc_kernel", "kernel_ms_avg": kern_ms_avg,
                "launches_per_step": launches_per_step, "algorithmic_flops_per_launch": alg_flops_launch,
                "executed_tflops": 3.0 * ach,
                "kernel_share_of_step": kern_ms_avg * launches_per_step / prof_ms_step if prof_ms_step > 0 else None}
    line = {
        "metric": metric_name(wl), "value": value, "unit": unit_name(wl), "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": wl["dtype"], "data": "synthetic",
        "config": {"workload": wl["desc"], "n_global": n, "d": d, "k": k, "rows_per_rank": n_loc,
                   "l2": "inputs larger than L2 (per-rank shard %.1f GB vs 126 MB)" % (alg_bytes_rank / 1e9),
                   "parallelism": f"split0 x{world}", "path": args.path, "dataset": args.data,
                   "comm": eng.comm_mode()},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roof,
    }
    if filt is not None:
        line["filter"] = filt
    if parity is not None:
        line["parity"] = parity
    if graph is not None:
        line["graph_replay"] = graph
    if world == 1 and not args.no_cpu:
        n_sample = SAMPLE_ROWS[args.workload]
        sec, thr, impl = timed_cpu_reference(n_sample, d, k, dtype, 3, 1, wl["kind"], args.data)
        scale = n / n_sample
        what = ("unmodified reference (heat from baseline/_ref under oracle/mpi4py_shim, torch CPU" if impl == "reference"
                else "oracle port of the reference (torch CPU")
        line["cpu_baseline"] = {
            "value": 1.0 / (sec * scale), "unit": unit_name(wl), "cores": thr, "kind": impl,
            "sample": (f"{what}, {thr} threads of {os.cpu_count()} host cores) on "
                       f"{n_sample} of {n} rows, {sec * 1e3:.1f} ms/step, extrapolated linearly in N (x{scale:.0f})")}
        if wl["kind"] == "kmeans" and args.ref_gpu_rows > 0:
            torch.cuda.empty_cache()
            rg = reference_on_gpu(args.ref_gpu_rows, d, k, dtype, dev, args.data)
            if rg is not None:
                if "sec_per_iter" in rg:
                    rg["iter_per_s_extrapolated"] = 1.0 / (rg["sec_per_iter"] * n / rg["rows"])
                line["reference_gpu"] = rg
    if world == 1 and args.workload == "config3" and args.data == "blobs" and not args.no_extras and not args.rows:
        line["other_workloads"] = run_extras(args)
    print(json.dumps(line), flush=True)


def run_extras(args):
    """The other BASELINE.json configurations (5: fp64 small k, 2: cdist, 4: large k) and the headline configuration
    on inputs that are NOT well-separated blobs, each measured by a short run of this same script in a child process
    (N=1 only) and folded into the headline line, so that the driver's single default run sees them all."""
    out = {}
    jobs = [("config5", "blobs"), ("config2", "blobs"), ("config4", "blobs"), ("config3", "randn"), ("config3", "uncentred")]
    for wl, data in jobs:
        key = wl if data == "blobs" else f"{wl}:{data}"
        cmd = [sys.executable, os.path.abspath(__file__), "--workload", wl, "--data", data, "--steps", "8", "--warmup", "3",
               "--no-e2e", "--no-cpu", "--no-extras", "--ref-gpu-rows", "0", "--path", args.path]
        try:
            res = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
            js = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
            if not js:
                out[key] = {"error": (res.stderr or res.stdout)[-200:]}
                continue
            j = json.loads(js[-1])
            r = j["roofline"]
            ent = {"metric": j["metric"], "value": j["value"], "unit": j["unit"], "ms_per_step": j["ms_per_step"],
                   "kernel": r["kernel"], "kernel_ms_avg": r["kernel_ms_avg"], "bound": r["bound"], "roofline_frac": r["frac"],
                   "steps": j["steps"], "workload": j["config"]["workload"]}
            if j.get("parity"):
                ent["parity_ok"] = j["parity"]["ok"]
            if j.get("filter"):
                ent["undecided_frac"] = j["filter"]["undecided_frac"]
            out[key] = ent
        except Exception as ex:  # reported, never hidden
            out[key] = {"error": str(ex)[:200]}
    # the "next" rows: rbf / manhattan tiles, KMedians / KMedoids passes, kNN — device times in ms
    try:
        res = subprocess.run([sys.executable, os.path.join(os.path.dirname(os.path.abspath(__file__)), "tools",
                                                           "consumers_bench.py")],
                             capture_output=True, text=True, timeout=240, cwd=os.path.dirname(os.path.abspath(__file__)))
        js = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
        out["next_rows_ms"] = json.loads(js[-1]) if js else {"error": (res.stderr or res.stdout)[-200:]}
    except Exception as ex:
        out["next_rows_ms"] = {"error": str(ex)[:200]}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="config3", choices=list(WORKLOADS))
    ap.add_argument("--path", default="auto", choices=["auto", "simt", "tc", "generic", "row128"])
    ap.add_argument("--data", default="blobs", choices=["blobs", "overlap", "randn", "uncentred"],
                    help="input distribution (blobs = the benchmark's; the others show the path off its best case)")
    ap.add_argument("--no-row-ws", action="store_true", help="run without the per-matrix |x| bound workspace")
    ap.add_argument("--ref-gpu-rows", type=int, default=4_000_000,
                    help="rows for the second bar: the unmodified reference on this GPU through torch CUDA (0 = skip)")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the short runs of the other configurations that the default N=1 run folds into its line")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--rows", type=int, default=None, help="override the global row count (experiments only)")
    args = ap.parse_args()
    wl = dict(WORKLOADS[args.workload])
    if args.rows:
        wl["n"] = args.rows
        wl["desc"] += f" [n overridden to {args.rows}]"
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_ours(args, wl)
    try:
        import torch.distributed as dist

        if dist.is_initialized():
            dist.destroy_process_group()
    except Exception:
        pass


if __name__ == "__main__":
    main()
