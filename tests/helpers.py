"""Shared test helpers: golden loading and the parity rule of BASELINE.json's north_star."""
from __future__ import annotations

import hashlib
import os

import numpy as np
import torch

from oracle import kmeans_oracle as orc

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# tolerances stated by the north_star: centroids 1e-5 relative (fp32) / 1e-12 (fp64)
CENTER_TOL = {torch.float32: 1e-5, torch.float64: 1e-12}


def load_golden(name: str):
    g = np.load(os.path.join(GOLD, f"{name}.npz"))
    return {k: g[k] for k in g.files}


def sha(t: torch.Tensor) -> str:
    return hashlib.sha256(t.contiguous().cpu().numpy().tobytes()).hexdigest()


def check_inputs(name, x, init, gold):
    assert sha(x) == str(gold["x_sha"]), f"{name}: regenerated X differs from the one the reference saw"
    assert sha(init) == str(gold["init_sha"]), f"{name}: regenerated init differs"


def assert_fit_matches(name, x, init, gold, centers, labels, n_iter, inertia, pre_centers=None):
    """n_iter equal; centroids within tolerance; labels exact except classified near-ties."""
    assert int(n_iter) == int(gold["n_iter"]), f"{name}: n_iter {n_iter} != reference {int(gold['n_iter'])}"
    ref_c = torch.from_numpy(gold["centers"])
    assert centers.dtype == ref_c.dtype, f"{name}: centroid dtype {centers.dtype} != {ref_c.dtype}"
    err = orc.centers_rel_err(ref_c, centers.cpu())
    tol = CENTER_TOL[ref_c.dtype]
    assert err <= tol, f"{name}: centroid relative error {err:.3e} > {tol}"
    ref_l = torch.from_numpy(gold["labels"].astype(np.int64))
    new_l = labels.cpu().view(-1).long()
    assert new_l.numel() == ref_l.numel()
    if pre_centers is None:
        assert torch.equal(ref_l, new_l), f"{name}: labels differ in {(ref_l != new_l).sum().item()} rows"
    else:
        par = orc.compare_labels(x, pre_centers.cpu(), ref_l, new_l)
        assert par.hard == 0, f"{name}: {par}"
    ref_in = float(gold["inertia"])
    got_in = float(inertia)
    assert abs(got_in - ref_in) <= 1e-3 * max(abs(ref_in), 1e-12) + 1e-10, f"{name}: inertia {got_in} vs {ref_in}"
