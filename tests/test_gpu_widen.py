"""GPU parity for the SURVEY.md section 8(f) rows built on top of the hot path:
RegularSpatial and KMedoids (8f-3), transition counting on the labels (8f-4) and
the device-resident tICA -> cluster pipeline (8f-1/2).  Checked against the
oracle restatements, the goldens written from the reference's own classes and the
known answers of msmbuilder/tests/test_transition_counts.py."""
import os
import pickle

import numpy as np
import pytest

from oracle import cluster_oracle as co
from oracle import libdistance_oracle as lo
from oracle import msm_oracle as mo
from oracle.gen_golden import cluster_inputs, msm_inputs

pytestmark = pytest.mark.gpu
MORE = [("euclidean", 4.0), ("cityblock", 8.0), ("chebyshev", 2.5)]


# ------------------------------------------------------------------ RegularSpatial
@pytest.mark.parametrize("metric,d_min", MORE)
@pytest.mark.parametrize("dtype", ["float32", "float64"])
def test_regular_spatial_matches_golden(golden_dir, metric, d_min, dtype):
    from msmbuilder_b200.cluster import RegularSpatial
    g = np.load(os.path.join(golden_dir, "cluster_more.npz"))
    seqs = cluster_inputs(31, 3, 300, 5, np.dtype(dtype))
    m = RegularSpatial(d_min=d_min, metric=metric).fit(seqs)
    key = "rs_%s_%s_" % (metric, dtype)
    np.testing.assert_array_equal(m.cluster_center_indices_, g[key + "ids"])
    np.testing.assert_array_equal(m.cluster_centers_, g[key + "centers"])
    assert m.n_clusters_ == len(g[key + "ids"])
    np.testing.assert_array_equal(np.concatenate(m.predict(seqs)), g[key + "predict"])
    m2 = pickle.loads(pickle.dumps(m))
    np.testing.assert_array_equal(np.concatenate(m2.predict(seqs)), g[key + "predict"])


@pytest.mark.parametrize("metric", list(lo.VECTOR_METRICS))
def test_regular_spatial_vs_oracle_all_metrics(metric):
    from msmbuilder_b200.cluster import RegularSpatial
    rs = np.random.RandomState(4)
    X = (rs.randn(3000, 8) * 2).astype(np.float32)
    if metric in ("hamming", "jaccard"):
        X = np.round(X)
    # choose d_min as a low quantile of the distances to frame 0 so that a few dozen centres appear
    d0 = lo.dist(X, X[0], metric)
    d_min = float(np.quantile(d0[1:], 0.35))
    ids, centers = co.regular_spatial_fit(X, d_min, metric)
    m = RegularSpatial(d_min=d_min, metric=metric).fit([X[:1000], X[1000:]])
    assert m.n_clusters_ == len(ids) and len(ids) > 1
    np.testing.assert_array_equal(m.cluster_center_indices_, co.split_indices(ids, [1000, 2000]))
    np.testing.assert_array_equal(m.cluster_centers_, centers)


def test_regular_spatial_every_frame_and_single_centre():
    from msmbuilder_b200.cluster import RegularSpatial
    X = np.arange(50, dtype=np.float64).reshape(-1, 1)
    m = RegularSpatial(d_min=0.5).fit([X])
    assert m.n_clusters_ == 50                     # every frame is farther than 0.5 from the others
    m = RegularSpatial(d_min=1e9).fit([X])
    assert m.n_clusters_ == 1 and list(m.cluster_center_indices_[0]) == [0, 0]
    m = RegularSpatial(d_min=2.0).fit([X])         # strict '>' (regularspatial.py:76): 0, 3, 6, ...
    np.testing.assert_array_equal(m.cluster_center_indices_[:, 1], np.arange(0, 50, 3))


def test_regular_spatial_skips_nan_frames_like_the_reference():
    # regularspatial.py:70-77: `np.all(d > d_min)` is False for a frame with a NaN coordinate, so such a
    # frame never becomes a centre (the running-minimum pass used to leave it at +inf)
    from msmbuilder_b200.cluster import RegularSpatial
    rs = np.random.RandomState(3)
    X = rs.randn(400, 5).astype(np.float32) * 3
    X[7, 2] = np.nan
    X[100] = np.nan
    want = [0]
    for i in range(1, len(X)):
        d = np.sqrt(((X[i].astype(np.float64) - X[want].astype(np.float64)) ** 2).sum(1))
        if np.all(d > 2.5):
            want.append(i)
    m = RegularSpatial(d_min=2.5, metric="euclidean").fit([X])
    assert 7 not in want and 100 not in want
    np.testing.assert_array_equal(m.cluster_centers_, X[want])


# ------------------------------------------------------------------ KMedoids
@pytest.mark.parametrize("metric,d_min", MORE)
@pytest.mark.parametrize("dtype", ["float32", "float64"])
def test_kmedoids_matches_golden(golden_dir, metric, d_min, dtype):
    from msmbuilder_b200.cluster import KMedoids
    g = np.load(os.path.join(golden_dir, "cluster_more.npz"))
    seqs = cluster_inputs(31, 3, 300, 5, np.dtype(dtype))
    for n_passes in (1, 4):
        km = KMedoids(n_clusters=5, n_passes=n_passes, metric=metric, random_state=7).fit(seqs)
        key = "km%d_%s_%s_" % (n_passes, metric, dtype)
        np.testing.assert_array_equal(km.cluster_ids_, g[key + "ids"])
        np.testing.assert_array_equal(np.concatenate(km.labels_), g[key + "labels"])
        # distances are reduced across lanes, not serially: last-bit differences in float64
        assert km.inertia_ == pytest.approx(float(g[key + "inertia"]), rel=1e-12)
        # a medoid is its own nearest centre
        pred = np.concatenate(km.predict(seqs))
        flat_ids = [300 * int(t) + int(f) for t, f in km.cluster_ids_]
        np.testing.assert_array_equal(pred[flat_ids], np.arange(5))


def test_kmedoids_argument_errors():
    from msmbuilder_b200.cluster import KMedoids
    X = [np.random.RandomState(0).randn(20, 2)]
    with pytest.raises(ValueError):
        KMedoids(n_clusters=2, n_passes=0).fit(X)
    with pytest.raises(ValueError):
        KMedoids(n_clusters=0).fit(X)
    with pytest.raises(ValueError):
        KMedoids(n_clusters=30).fit(X)


# ------------------------------------------------------------------ transition counts
def test_transition_counts_known_answers():
    # msmbuilder/tests/test_transition_counts.py:7-26,38-60,66-78; core.py:517-532
    from msmbuilder_b200.msm import transition_counts as tc
    with pytest.raises(ValueError):
        tc([1, 2, 3])
    c, m = tc([np.arange(10)])
    np.testing.assert_array_equal(c, np.eye(10, k=1))
    assert list(m.keys()) == list(range(10)) and list(m.values()) == list(range(10))
    c, m = tc([range(10)], lag_time=2)
    np.testing.assert_array_equal(c, 0.5 * np.eye(10, k=2))
    c, m = tc([[100000000, 100000000, 100000001, 100000001]])
    np.testing.assert_array_equal(c, np.array([[1., 1.], [0., 1.]]))
    assert m == {100000000: 0, 100000001: 1}
    c, m = tc([[0, 0, 0, 1, 1]])
    np.testing.assert_array_equal(c, np.array([[2., 1.], [0., 1.]]))
    c, m = tc([[100, 200, 300]])
    np.testing.assert_array_equal(c, np.eye(3, k=1))
    assert m == {100: 0, 200: 1, 300: 2}
    c, m = tc([[0]])
    np.testing.assert_array_equal(c, np.zeros((1, 1)))
    c, m = tc([[0, np.nan]])
    assert m == {0: 0}
    np.testing.assert_array_equal(c, np.zeros((1, 1)))
    c, m = tc([[np.nan]])
    assert m == {}
    np.testing.assert_array_equal(c, np.zeros((0, 0)))
    C, _ = tc([np.arange(6)], lag_time=3)
    np.testing.assert_array_almost_equal(C, np.eye(6, k=3) / 3)
    X = np.arange(10)
    C1, m1 = tc([X], lag_time=3, sliding_window=False)
    C2, m2 = tc([X[::3]], sliding_window=True)
    np.testing.assert_array_almost_equal(C1, C2)
    assert m1 == m2


def test_transition_counts_match_golden(golden_dir):
    from msmbuilder_b200.msm import transition_counts as tc
    g = np.load(os.path.join(golden_dir, "msm_counts.npz"))
    for case in range(int(g["n_cases"])):
        seqs = msm_inputs(case)
        np.testing.assert_array_equal(np.concatenate(seqs), g["labels_%d" % case])
        c, m = tc(seqs, lag_time=int(g["lag_%d" % case]), sliding_window=bool(g["sliding_%d" % case]))
        np.testing.assert_array_equal(c, g["counts_%d" % case])
        np.testing.assert_array_equal(np.array(sorted(m.keys())), g["classes_%d" % case])


@pytest.mark.parametrize("n_states,lag,sliding", [(5, 1, True), (90, 4, True), (91, 3, False),
                                                  (700, 10, True), (2000, 2, True)])
def test_transition_counts_vs_oracle_large(n_states, lag, sliding):
    import torch
    from msmbuilder_b200.msm import transition_counts as tc
    rs = np.random.RandomState(n_states)
    seqs = []
    for n in (200000, 1, lag, lag + 1, 77777, 300001):
        # sticky chains: long runs of one state = the contended-bin case
        y = rs.randint(0, n_states, size=n)
        keep = rs.rand(n) < 0.9
        for i in range(1, n):
            if keep[i]:
                y[i] = y[i - 1]
        seqs.append(y.astype(np.int64))
    c_ref, m_ref = mo.transition_counts(seqs, lag, sliding)
    c, m = tc(seqs, lag, sliding)
    np.testing.assert_array_equal(c, c_ref)
    assert m == m_ref
    # int32 CUDA tensors (what the assignment kernels leave on the device) give the same
    dev_seqs = [torch.from_numpy(s.astype(np.int32)).cuda() for s in seqs]
    c2, m2 = tc(dev_seqs, lag, sliding)
    np.testing.assert_array_equal(c2, c_ref)
    assert m2 == m_ref
    # conservation: every in-sequence pair is counted once
    step = 1 if (sliding or lag == 1) else lag
    n_pairs = sum(len(range(0, max(len(s) - lag, 0), step)) for s in seqs)
    assert c.sum() * (lag if step == 1 else 1) == pytest.approx(n_pairs, abs=1e-6)


def test_transition_counts_float_labels_with_missing():
    from msmbuilder_b200.msm import transition_counts as tc
    rs = np.random.RandomState(8)
    seqs = []
    for n in (5000, 1, 12345):
        y = rs.randint(0, 7, size=n).astype(np.float64)
        y[rs.rand(n) < 0.05] = np.nan
        seqs.append(y)
    c_ref, m_ref = mo.transition_counts(seqs, 3)
    c, m = tc(seqs, 3)
    np.testing.assert_array_equal(c, c_ref)
    assert m == m_ref
    with pytest.raises(TypeError):
        tc([np.array([0.5, 1.0])])
    with pytest.raises(TypeError):
        tc([np.array(["a", "b"])])


# ------------------------------------------------------------------ device-resident pipeline
def test_upload_once_pipeline_equals_host_calls():
    import torch
    import msmbuilder_b200 as mb
    from msmbuilder_b200.decomposition import tICA
    from msmbuilder_b200.cluster import KCenters
    from msmbuilder_b200.msm import transition_counts as tc
    from msmbuilder_b200.synthetic import ar1_numpy
    host = ar1_numpy(4, 5000, 64, seed=3, dtype=np.float32)
    dseqs = mb.device_sequences(host)
    assert all(t.is_cuda for t in dseqs)
    base = dseqs[0].data_ptr()
    assert dseqs[1].data_ptr() == base + 5000 * 64 * 4          # back to back: adopted, not copied

    t_host = tICA(n_components=3, lag_time=5).fit(host)
    t_dev = tICA(n_components=3, lag_time=5).fit(dseqs)
    np.testing.assert_array_equal(t_host.eigenvalues_, t_dev.eigenvalues_)

    tics_dev = t_dev.transform(dseqs)                            # stays on the device
    assert all(y.is_cuda and y.dtype == torch.float64 for y in tics_dev)
    tics_host = t_host.transform(host)
    for a, b in zip(tics_dev, tics_host):
        np.testing.assert_array_equal(a.cpu().numpy(), b)

    kc_dev = KCenters(n_clusters=6, random_state=1).fit(tics_dev)
    kc_host = KCenters(n_clusters=6, random_state=1).fit(tics_host)
    assert kc_dev.cluster_ids_ == kc_host.cluster_ids_
    for a, b in zip(kc_dev.labels_, kc_host.labels_):
        np.testing.assert_array_equal(a, b)
    c_dev, _ = tc(kc_dev.labels_, lag_time=5)
    c_ref, _ = mo.transition_counts(kc_host.labels_, 5)
    np.testing.assert_array_equal(c_dev, c_ref)


# ------------------------------------------------------------------ .npy directory stream (8f-2)
def _write_dataset(tmp_path, dtype=np.float32):
    from msmbuilder_b200.io import save_sequences
    from msmbuilder_b200.synthetic import ar1_numpy
    seqs = ar1_numpy(5, 3000, 32, seed=9, dtype=dtype)
    seqs = [seqs[0], seqs[1][:7], seqs[2][:1234], seqs[3], seqs[4][:2999]]   # ragged, one too short
    return seqs, save_sequences(str(tmp_path / "ds"), seqs)


def test_npy_stream_feeds_tica_and_clusterers(tmp_path):
    import warnings
    import torch
    from msmbuilder_b200.io import NumpyDirStream
    from msmbuilder_b200.decomposition import tICA
    from msmbuilder_b200.cluster import KCenters
    seqs, path = _write_dataset(tmp_path)
    stream = NumpyDirStream(path, prefetch=2)
    assert len(stream) == 5 and [s for s, _ in stream.shapes()] == [x.shape for x in seqs]

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")                      # the 7-frame sequence is skipped
        a = tICA(n_components=3, lag_time=10).fit(seqs)
        b = tICA(n_components=3, lag_time=10).fit(stream)
    assert a.n_sequences_ == b.n_sequences_ == 4 and a.n_observations_ == b.n_observations_
    # same kernel on the same frames; only the host-side order of the float64 folds differs
    np.testing.assert_allclose(a.eigenvalues_, b.eigenvalues_, rtol=0, atol=1e-12)
    scale = np.abs(a._outer_0_to_T_lagged).max()
    np.testing.assert_allclose(a._outer_0_to_T_lagged, b._outer_0_to_T_lagged, rtol=0, atol=1e-12 * scale)

    # iterating twice gives the same tensors; an abandoned iteration does not hang
    first = [t.cpu().numpy() for t in stream]
    for x, y in zip(first, seqs):
        np.testing.assert_array_equal(x, y)
    for i, t in enumerate(stream):
        assert t.is_cuda
        if i == 1:
            break

    dseqs = stream.to_device()
    assert dseqs[1].data_ptr() == dseqs[0].data_ptr() + seqs[0].nbytes
    for x, y in zip(dseqs, seqs):
        np.testing.assert_array_equal(x.cpu().numpy(), y)

    ka = KCenters(n_clusters=7, random_state=2).fit(seqs)
    kb = KCenters(n_clusters=7, random_state=2).fit(stream)
    assert ka.cluster_ids_ == kb.cluster_ids_
    for x, y in zip(ka.labels_, kb.labels_):
        np.testing.assert_array_equal(x, y)
    for x, y in zip(ka.predict(seqs), kb.predict(stream)):
        np.testing.assert_array_equal(x, y)
    for x, y in zip(a.transform(seqs), b.transform(stream)):      # a, b: equal up to float64 atomics order
        np.testing.assert_allclose(x, y.cpu().numpy(), rtol=0, atol=1e-10)
    for x, y in zip(b.transform(seqs), b.transform(stream)):      # same model: same bits
        np.testing.assert_array_equal(x, y.cpu().numpy())


def test_npy_stream_small_stage_batches(tmp_path):
    # force one accumulate call per sequence: results must not depend on batching
    import warnings
    from msmbuilder_b200.io import NumpyDirStream
    from msmbuilder_b200.decomposition import tICA
    seqs, path = _write_dataset(tmp_path, np.float64)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        a = tICA(n_components=2, lag_time=3).fit(seqs)
        b = tICA(n_components=2, lag_time=3)
        b._stage_bytes = 1
        b.fit(NumpyDirStream(path, prefetch=1))
    np.testing.assert_allclose(a.eigenvalues_, b.eigenvalues_, rtol=0, atol=1e-12)
    assert a.n_observations_ == b.n_observations_
