"""CPU: pin the clustering-loop oracle (oracle/cluster_oracle.py) against the
reference's own kcenters.py / minibatchkmedoids.py (verbatim, build container) and
against tests/golden/cluster_small.npz; plus the reference's known-answer tests
msmbuilder/tests/test_kcenters.py:29-106 and tests/test_kmedoids.py:53-91."""
import os

import numpy as np
import pytest
import scipy.spatial.distance

from oracle import cluster_oracle as co
from oracle import libdistance_oracle as lo
from oracle import ref_loader
from oracle.gen_golden import cluster_inputs

needs_ref = pytest.mark.skipif(not ref_loader.available(), reason="/root/reference absent")
METRICS = list(lo.VECTOR_METRICS)


def _inputs(metric, dtype):
    seqs = cluster_inputs(21, 3, 400, 5, np.dtype(dtype))
    if metric in ("hamming", "jaccard"):
        seqs = [np.round(s).astype(dtype) for s in seqs]
    return seqs


@pytest.mark.parametrize("metric", METRICS)
@pytest.mark.parametrize("dtype", ["float32", "float64"])
def test_oracle_matches_golden(golden_dir, metric, dtype):
    g = np.load(os.path.join(golden_dir, "cluster_small.npz"))
    seqs = _inputs(metric, dtype)
    X = np.concatenate(seqs)
    r = co.kcenters_fit(X, 9, metric, random_state=3)
    key = "kc_%s_%s_" % (metric, dtype)
    np.testing.assert_array_equal(r["cluster_ids_"], g[key + "ids"])
    np.testing.assert_array_equal(r["labels_"], g[key + "labels"])
    np.testing.assert_array_equal(r["distances_"], g[key + "distances"])
    assert r["inertia_"] == float(g[key + "inertia"])
    pred, _ = lo.assign_nearest(X, r["cluster_centers_"], metric)
    np.testing.assert_array_equal(pred, g[key + "predict"])
    m = co.minibatch_kmedoids_fit(X, 6, max_iter=3, batch_size=40, metric=metric, random_state=5)
    key = "mb_%s_%s_" % (metric, dtype)
    np.testing.assert_array_equal(co.split_indices(m["cluster_ids_"], [400] * 3), g[key + "ids"])
    np.testing.assert_array_equal(m["labels_"], g[key + "labels"])
    assert m["inertia_"] == float(g[key + "inertia"])


@needs_ref
def test_oracle_matches_verbatim_reference():
    KCenters, MiniBatchKMedoids, _ = ref_loader.load_cluster()
    rs = np.random.RandomState(0)
    seqs = [rs.randn(150, 4).astype(np.float32) for _ in range(4)]
    X = np.concatenate(seqs)
    for metric in ("euclidean", "cityblock", "canberra"):
        kc = KCenters(n_clusters=11, metric=metric, random_state=7).fit(seqs)
        r = co.kcenters_fit(X, 11, metric, random_state=7)
        assert kc.cluster_ids_ == r["cluster_ids_"]
        np.testing.assert_array_equal(np.concatenate(kc.labels_), r["labels_"])
        np.testing.assert_array_equal(np.concatenate(kc.distances_), r["distances_"])
        mb = MiniBatchKMedoids(n_clusters=5, batch_size=30, metric=metric, random_state=9).fit(seqs)
        m = co.minibatch_kmedoids_fit(X, 5, batch_size=30, metric=metric, random_state=9)
        np.testing.assert_array_equal(mb.cluster_ids_, co.split_indices(m["cluster_ids_"], [150] * 4))
        np.testing.assert_array_equal(np.concatenate(mb.labels_), m["labels_"])
        assert mb.inertia_ == m["inertia_"]


def test_three_clusters():
    # test_kcenters.py:29-44: three point masses, k=2, random_state=0
    X = np.concatenate([np.zeros((10, 2)), np.ones((10, 2)), 0.5 * np.ones((5, 2))]).astype(np.float64)
    r = co.kcenters_fit(X, 2, "euclidean", random_state=0)
    cs = {tuple(c) for c in r["cluster_centers_"]}
    assert cs == {(0.0, 0.0), (1.0, 1.0)}
    assert set(np.round(r["distances_"], 8)) <= {0.0, round(np.sqrt(2) / 2, 8)}


def test_fit_predict_equals_cdist_argmin():
    # test_kcenters.py:47-71
    X = np.random.RandomState(0).randn(300, 3).astype(np.float32)
    for metric in ("euclidean", "cityblock"):
        r = co.kcenters_fit(X, 10, metric, random_state=0)
        D = scipy.spatial.distance.cdist(X.astype(np.float64), r["cluster_centers_"].astype(np.float64), metric=metric)
        np.testing.assert_array_equal(r["labels_"], D.argmin(1))


def test_sqeuclidean_same_labels_and_dtype_agreement():
    # test_kcenters.py:74-106
    X = np.random.RandomState(1).randn(200, 4)
    a = co.kcenters_fit(X, 8, "euclidean", random_state=0)
    b = co.kcenters_fit(X, 8, "sqeuclidean", random_state=0)
    np.testing.assert_array_equal(a["labels_"], b["labels_"])
    X32 = X.astype(np.float32)
    c = co.kcenters_fit(X32, 8, "euclidean", random_state=0)
    d = co.kcenters_fit(X32.astype(np.float64), 8, "euclidean", random_state=0)
    np.testing.assert_array_equal(c["labels_"], d["labels_"])
    np.testing.assert_allclose(c["distances_"], d["distances_"], rtol=1e-6)
