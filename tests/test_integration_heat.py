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
        # the other consumers are patched too and fall through on CPU arrays with the reference's own results
        from cases import consumer_inputs

        g, inp = load_golden("consumers"), consumer_inputs()
        assert "fit" in ht.cluster.KMedoids.__dict__ and "_assign_to_cluster" in ht.cluster.KMedians.__dict__
        hx, hinit = ht.array(inp["x"], split=0), ht.array(inp["init"])
        kmed = ht.cluster.KMedians(n_clusters=4, init=hinit, max_iter=30, tol=1e-4).fit(hx)
        assert kmed.n_iter_ == int(g["kmedians_f32_n_iter"])
        assert np.array_equal(kmed.cluster_centers_.larray.numpy(), g["kmedians_f32_centers"])
        assert np.array_equal(kmed.predict(hx).larray.numpy(), g["kmedians_f32_predict"])
        bp = ht.cluster.BatchParallelKMeans(n_clusters=4, init="k-means++", max_iter=30, tol=1e-4, random_state=5).fit(hx)
        assert np.array_equal(bp.cluster_centers_.larray.numpy(), g["bpkmeans_f32_centers"])
        assert np.array_equal(bp.predict(hx).larray.numpy(), g["bpkmeans_f32_predict"])
        knn = ht.classification.kneighborsclassifier.KNeighborsClassifier(n_neighbors=5)
        knn.fit(hx, ht.array(inp["y"], split=0))
        assert np.array_equal(knn.predict(ht.array(inp["x_test"], split=0)).larray.numpy(), g["knn_f32_classes"])
    finally:
        hki.uninstall()
    assert ht.cluster.KMeans.fit is orig_fit
    assert "fit" not in ht.cluster.KMedoids.__dict__ or ht.cluster.KMedoids.fit.__name__ == "fit"
    assert "_assign_to_cluster" not in ht.cluster.KMedians.__dict__  # inherited from _KCluster again
