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


# ----------------------------------------------------------------- SURVEY 8f-3 consumers
MORE = [("euclidean", 4.0), ("cityblock", 8.0), ("chebyshev", 2.5)]


@pytest.mark.parametrize("metric,d_min", MORE)
@pytest.mark.parametrize("dtype", ["float32", "float64"])
def test_regular_spatial_and_kmedoids_oracle_match_golden(golden_dir, metric, d_min, dtype):
    g = np.load(os.path.join(golden_dir, "cluster_more.npz"))
    seqs = cluster_inputs(31, 3, 300, 5, np.dtype(dtype))
    X = np.concatenate(seqs)
    ids, centers = co.regular_spatial_fit(X, d_min, metric)
    key = "rs_%s_%s_" % (metric, dtype)
    np.testing.assert_array_equal(co.split_indices(ids, [300] * 3), g[key + "ids"])
    np.testing.assert_array_equal(centers, g[key + "centers"])
    pred, _ = lo.assign_nearest(X, centers, metric)
    np.testing.assert_array_equal(pred, g[key + "predict"])
    for n_passes in (1, 4):
        r = co.kmedoids_fit(X, 5, n_passes, metric, random_state=7)
        key = "km%d_%s_%s_" % (n_passes, metric, dtype)
        np.testing.assert_array_equal(co.split_indices(r["cluster_ids_"], [300] * 3), g[key + "ids"])
        np.testing.assert_array_equal(r["labels_"], g[key + "labels"])
        assert r["inertia_"] == float(g[key + "inertia"])


@needs_ref
def test_kmedoids_restarts_port_matches_reference_cxx():
    # the restart loop of kmedoids.cc:160-250 incl. its RandomState use (:314-383)
    rs = np.random.RandomState(5)
    for n, k, n_pass in ((30, 3, 1), (80, 6, 2), (120, 9, 7), (12, 12, 3), (9, 1, 2)):
        X = rs.randn(n, 3)
        dm = lo.pdist(X, "euclidean", impl="reference")
        a = lo.kmedoids(k, dm, n_pass, random_state=11, impl="reference")
        b = lo.kmedoids(k, dm, n_pass, random_state=11, impl="port")
        np.testing.assert_array_equal(a[0], b[0])
        assert a[1] == b[1] and a[2] == b[2]


@needs_ref
def test_more_clusterers_match_verbatim_reference():
    RegularSpatial, KMedoids = ref_loader.load_more_clusterers()
    rs = np.random.RandomState(2)
    seqs = [rs.randn(120, 3).astype(np.float32) for _ in range(3)]
    X = np.concatenate(seqs)
    ref = RegularSpatial(d_min=1.2, metric="euclidean").fit(seqs)
    ids, centers = co.regular_spatial_fit(X, 1.2, "euclidean")
    np.testing.assert_array_equal(ref.cluster_center_indices_, co.split_indices(ids, [120] * 3))
    assert ref.n_clusters_ == len(ids)
    km = KMedoids(n_clusters=4, n_passes=3, random_state=0).fit(seqs)
    r = co.kmedoids_fit(X, 4, 3, "euclidean", random_state=0)
    np.testing.assert_array_equal(km.cluster_ids_, co.split_indices(r["cluster_ids_"], [120] * 3))
    np.testing.assert_array_equal(np.concatenate(km.labels_), r["labels_"])
    assert km.inertia_ == r["inertia_"]
