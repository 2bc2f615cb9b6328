"""The Heat binding (heat_b200/integration.py) against the real reference — only where /root/reference exists
(the build container): patched classes must fall through to the reference code for CPU arrays and give
bit-identical results; on a GPU box without Heat the test is skipped."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present")
def test_install_patches_and_defers_on_cpu():
    sys.path.insert(0, os.path.join(ROOT, "oracle", "mpi4py_shim"))
    sys.path.insert(1, REF)
    import heat as ht

    import heat_b200.integration as hki
    from cases import make_case
    from helpers import load_golden

    orig_fit = ht.cluster.KMeans.fit
    assert hki.install() is True
    assert ht.cluster.KMeans.fit is not orig_fit
    try:
        x, init = make_case("blobs_f32_d8_k6")
        gold = load_golden("blobs_f32_d8_k6")
        km = ht.cluster.KMeans(n_clusters=6, init=ht.array(init), max_iter=300, tol=1e-4).fit(ht.array(x, split=0))
        assert km.n_iter_ == int(gold["n_iter"])
        assert torch.equal(km.cluster_centers_.larray, torch.from_numpy(gold["centers"]))
        assert np.array_equal(km.labels_.larray.view(-1).numpy(), gold["labels"].astype(np.int64))
        pred = km.predict(ht.array(x, split=0))
        assert np.array_equal(pred.larray.view(-1).numpy(), gold["predict_labels"].astype(np.int64))
        d = ht.spatial.cdist(ht.ones((4, 4), split=0), ht.zeros((6, 4)), quadratic_expansion=True)
        assert torch.equal(d.larray, torch.full((4, 6), 2.0))
    finally:
        hki.uninstall()
    assert ht.cluster.KMeans.fit is orig_fit
