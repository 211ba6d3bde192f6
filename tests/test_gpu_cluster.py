"""GPU parity: KCenters / MiniBatchKMedoids / MiniBatchKMeans estimators vs the
oracle loops, the committed goldens (written from the reference's own classes) and
the reference's known-answer tests (msmbuilder/tests/test_kcenters.py,
tests/test_kmedoids.py:53-132, tests/test_clustering.py:64-78)."""
import os
import pickle

import numpy as np
import pytest
import scipy.spatial.distance

from oracle import cluster_oracle as co
from oracle import libdistance_oracle as lo
from oracle.gen_golden import cluster_inputs

pytestmark = pytest.mark.gpu
METRICS = list(lo.VECTOR_METRICS)


def _inputs(metric, dtype):
    seqs = cluster_inputs(21, 3, 400, 5, np.dtype(dtype))
    if metric in ("hamming", "jaccard"):
        seqs = [np.round(s).astype(dtype) for s in seqs]
    return seqs


@pytest.mark.parametrize("metric", METRICS)
@pytest.mark.parametrize("dtype", ["float32", "float64"])
def test_kcenters_matches_golden(golden_dir, metric, dtype):
    from msmbuilder_b200.cluster import KCenters
    g = np.load(os.path.join(golden_dir, "cluster_small.npz"))
    seqs = _inputs(metric, dtype)
    kc = KCenters(n_clusters=9, metric=metric, random_state=3).fit(seqs)
    key = "kc_%s_%s_" % (metric, dtype)
    np.testing.assert_array_equal(kc.cluster_ids_, g[key + "ids"])
    np.testing.assert_array_equal(np.concatenate(kc.labels_), g[key + "labels"])
    np.testing.assert_allclose(np.concatenate(kc.distances_), g[key + "distances"], rtol=1e-13)
    np.testing.assert_allclose(kc.inertia_, float(g[key + "inertia"]), rtol=1e-12)
    np.testing.assert_array_equal(np.concatenate(kc.predict(seqs)), g[key + "predict"])
    assert [len(l) for l in kc.labels_] == [400, 400, 400]
    assert kc.cluster_centers_.dtype == np.dtype(dtype) and kc.cluster_centers_.shape == (9, 5)


@pytest.mark.parametrize("metric", METRICS)
@pytest.mark.parametrize("dtype", ["float32", "float64"])
def test_minibatch_kmedoids_matches_golden(golden_dir, metric, dtype):
    from msmbuilder_b200.cluster import MiniBatchKMedoids
    g = np.load(os.path.join(golden_dir, "cluster_small.npz"))
    seqs = _inputs(metric, dtype)
    mb = MiniBatchKMedoids(n_clusters=6, batch_size=40, max_iter=3, metric=metric, random_state=5).fit(seqs)
    key = "mb_%s_%s_" % (metric, dtype)
    np.testing.assert_array_equal(mb.cluster_ids_, g[key + "ids"])
    np.testing.assert_array_equal(np.concatenate(mb.labels_), g[key + "labels"])
    np.testing.assert_allclose(mb.inertia_, float(g[key + "inertia"]), rtol=1e-12)


def test_three_clusters():
    # test_kcenters.py:29-44
    from msmbuilder_b200.cluster import KCenters
    data = [np.zeros((10, 2)), np.ones((10, 2)), 0.5 * np.ones((5, 2))]
    m = KCenters(n_clusters=2, random_state=0).fit(data)
    assert {tuple(c) for c in m.cluster_centers_} == {(0.0, 0.0), (1.0, 1.0)}
    for d in m.distances_:
        assert set(np.round(d, 8)) <= {0.0, round(np.sqrt(2) / 2, 8)}


def test_shapes_fit_predict_pickle():
    # test_kcenters.py:11-26,82-92; test_clustering.py:64-78
    from msmbuilder_b200.cluster import KCenters
    rs = np.random.RandomState(0)
    data = [rs.randn(100, 3), rs.randn(50, 3), rs.randn(7, 3)]
    m = KCenters(n_clusters=4, random_state=1)
    labels = m.fit_predict(data)
    assert [l.shape for l in labels] == [(100,), (50,), (7,)]
    assert all(not np.isnan(l).any() for l in m.distances_)
    for a, b in zip(m.predict(data), labels):
        np.testing.assert_array_equal(a, b)
        assert a.dtype == np.intp
    np.testing.assert_array_equal(m.partial_predict(data[1]), labels[1])
    np.testing.assert_array_equal(m.transform(data)[0], labels[0])
    m2 = pickle.loads(pickle.dumps(m))
    np.testing.assert_array_equal(m2.predict(data)[0], labels[0])
    assert isinstance(m.summarize(), str)
    # labels == scipy cdist argmin (test_kcenters.py:47-71)
    D = scipy.spatial.distance.cdist(np.concatenate(data), m.cluster_centers_)
    np.testing.assert_array_equal(np.concatenate(labels), D.argmin(1))


def test_kcenters_large_random_vs_oracle():
    from msmbuilder_b200.cluster import KCenters
    rs = np.random.RandomState(3)
    X = rs.randn(20000, 16).astype(np.float32)
    r = co.kcenters_fit(X, 25, "euclidean", random_state=4)
    m = KCenters(n_clusters=25, random_state=4).fit([X[:7000], X[7000:]])
    assert m.cluster_ids_ == r["cluster_ids_"]
    np.testing.assert_array_equal(np.concatenate(m.labels_), r["labels_"])
    np.testing.assert_allclose(np.concatenate(m.distances_), r["distances_"], rtol=1e-13)


def test_kcenters_accepts_device_tensors_without_copy():
    import torch
    from msmbuilder_b200.cluster import KCenters
    rs = np.random.RandomState(5)
    X = rs.randn(3000, 8).astype(np.float32)
    Xd = torch.from_numpy(X).cuda()
    seqs = [Xd[:1000], Xd[1000:]]
    a = KCenters(n_clusters=6, random_state=2).fit(seqs)
    b = KCenters(n_clusters=6, random_state=2).fit([X[:1000], X[1000:]])
    assert a.cluster_ids_ == b.cluster_ids_
    np.testing.assert_array_equal(np.concatenate(a.labels_), np.concatenate(b.labels_))


def test_minibatch_kmedoids_reference_tests():
    # test_kmedoids.py:53-91,114-132
    from msmbuilder_b200.cluster import MiniBatchKMedoids
    rs = np.random.RandomState(0)
    X = [np.concatenate([rs.randn(30, 2) * 0.1, rs.randn(30, 2) * 0.1 + 10])]
    m = MiniBatchKMedoids(n_clusters=2, batch_size=10, random_state=0).fit(X)
    lab = m.labels_[0]
    assert len(set(lab[:30])) == 1 and len(set(lab[30:])) == 1 and lab[0] != lab[59]
    D = scipy.spatial.distance.cdist(X[0], m.cluster_centers_)
    np.testing.assert_allclose(m.inertia_, D.min(1).sum(), rtol=1e-10)
    assert m.cluster_ids_.shape == (2, 2)           # (traj_i, frame_i) pairs
    np.testing.assert_array_equal(X[0][m.cluster_ids_[:, 1]], m.cluster_centers_)
    with pytest.raises(ValueError):
        MiniBatchKMedoids(n_clusters=2, metric="blah").fit(X)


def test_minibatch_kmeans_labels_match_sklearn():
    # BASELINE config 3 (scaled down): labels = Euclidean argmin to the fitted centres
    from msmbuilder_b200.cluster import MiniBatchKMeans
    rs = np.random.RandomState(1)
    X = [rs.randn(20000, 16).astype(np.float32) * np.linspace(3, 0.3, 16).astype(np.float32)]
    m = MiniBatchKMeans(n_clusters=50, random_state=0, n_init=1).fit(X)
    D = scipy.spatial.distance.cdist(X[0].astype(np.float64), m.cluster_centers_.astype(np.float64), "sqeuclidean")
    Ds = np.sort(D, axis=1)
    clear = (Ds[:, 1] - Ds[:, 0]) > 1e-9 * Ds[:, 0]
    np.testing.assert_array_equal(m.labels_[0][clear], D.argmin(1)[clear])
    np.testing.assert_array_equal(m.predict(X)[0], m.labels_[0])
    from sklearn.cluster import MiniBatchKMeans as Sk
    sk = Sk(n_clusters=50, random_state=0, n_init=1).fit(X[0])
    np.testing.assert_allclose(sk.cluster_centers_, m.cluster_centers_)
