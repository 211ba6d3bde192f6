"""GPU parity: msmbuilder_b200.libdistance (through the C ABI) vs the CPU oracle,
every metric x dtype, the shapes/edge cases of msmbuilder/tests/test_libdistance.py."""
import numpy as np
import pytest

from oracle import libdistance_oracle as lo

pytestmark = pytest.mark.gpu
METRICS = list(lo.VECTOR_METRICS)


def _data(seed, n, d, dtype, metric):
    rs = np.random.RandomState(seed)
    X = rs.randn(n, d)
    if metric in ("hamming", "jaccard"):
        X = np.round(X)
    X[rs.rand(n, d) > 0.9] = 0.0
    return X.astype(dtype)


# distances: float64 accumulation on both sides, different summation order
RTOL = 1e-13


@pytest.mark.parametrize("metric", METRICS)
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("d", [1, 3, 16, 37, 64, 256])
def test_dist_cdist_pdist(metric, dtype, d):
    from msmbuilder_b200 import libdistance as ld
    X = _data(0, 203, d, dtype, metric)
    Y = _data(1, 11, d, dtype, metric)
    idx = np.random.RandomState(2).randint(0, 203, 29)
    np.testing.assert_allclose(ld.cdist(X, Y, metric), lo.cdist(X, Y, metric), rtol=RTOL, atol=1e-300, equal_nan=True)
    np.testing.assert_allclose(ld.dist(X, Y[3], metric), lo.dist(X, Y[3], metric), rtol=RTOL, equal_nan=True)
    np.testing.assert_allclose(ld.dist(X, Y[3], metric, idx), lo.dist(X, Y[3], metric, idx), rtol=RTOL, equal_nan=True)
    np.testing.assert_allclose(ld.pdist(X[:60], metric), lo.pdist(X[:60], metric), rtol=RTOL, equal_nan=True)
    np.testing.assert_allclose(ld.pdist(X, metric, idx), lo.pdist(X, metric, idx), rtol=RTOL, equal_nan=True)
    pairs = np.random.RandomState(3).randint(0, 203, (40, 2))
    a, b = ld.sumdist(X, metric, pairs), lo.sumdist(X, metric, pairs)
    assert (np.isnan(a) and np.isnan(b)) or abs(a - b) <= 1e-12 * max(1.0, abs(b))


@pytest.mark.parametrize("metric", METRICS)
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_assign_nearest(metric, dtype):
    from msmbuilder_b200 import libdistance as ld
    X = _data(4, 1000, 7, dtype, metric)
    Y = _data(5, 13, 7, dtype, metric)
    Y[5] = Y[2]                           # exact tie: lowest index must win
    idx = np.random.RandomState(6).randint(0, 1000, 77)
    for rows in (None, idx):
        labels, inertia = ld.assign_nearest(X, Y, metric, rows)
        ref_labels, ref_inertia = lo.assign_nearest(X, Y, metric, rows)
        D = lo.cdist(X if rows is None else X[rows], Y, metric)
        # labels are bit-exact wherever the oracle's two best distances differ by > 1e-12 relative
        Ds = np.sort(np.where(np.isnan(D), np.inf, D), axis=1)
        clear = (Ds[:, 1] - Ds[:, 0]) > 1e-12 * np.maximum(Ds[:, 0], 1e-300)
        tie = Ds[:, 1] == Ds[:, 0]
        np.testing.assert_array_equal(labels[clear | tie], ref_labels[clear | tie])
        assert labels.dtype == np.intp
        if np.isfinite(ref_inertia):
            assert abs(inertia - ref_inertia) <= 1e-11 * abs(ref_inertia)


def test_errors_match_reference():
    from msmbuilder_b200 import libdistance as ld
    X = np.zeros((4, 3), dtype=np.float32)
    with pytest.raises(ValueError):
        ld.cdist(X, X, "nope")                         # libdistance.pyx:122-124
    with pytest.raises(TypeError):
        ld.cdist(X, X.astype(np.float64), "euclidean")  # libdistance.pyx:130-131


def test_large_rows_and_ragged_widths():
    from msmbuilder_b200 import libdistance as ld
    rs = np.random.RandomState(7)
    for d in (2, 5, 130, 300):
        X = rs.randn(5000, d).astype(np.float32)
        y = rs.randn(d).astype(np.float32)
        np.testing.assert_allclose(ld.dist(X, y, "euclidean"), lo.dist(X, y, "euclidean"), rtol=RTOL)


@pytest.mark.parametrize("n,d,k", [(20000, 16, 500), (9000, 64, 100), (6000, 12, 37), (5000, 128, 180),
                                   (4100, 256, 33), (12000, 8, 2)])
def test_tensor_core_filter_labels_equal_exact_engine(n, d, k, monkeypatch):
    # K3 on tcgen05 (assign_umma.cu, n >= 4096 frames without a row gather): labels must be those of the
    # float64 scan (MSMB200_ASSIGN_EXACT) -- ties, duplicates and badly scaled features included
    from msmbuilder_b200 import libdistance as ld
    rs = np.random.RandomState(n + d + k)
    X = (rs.randn(n, d) * 10.0 ** rs.uniform(-2, 2, size=d)).astype(np.float32)
    Y = X[rs.choice(n, k, replace=False)].copy()
    Y[k // 2] = Y[0]                                   # duplicate centre: lowest index wins
    X[::7] = np.round(X[::7])                          # lattice points: exact ties between centres
    Y[1::5] = np.round(Y[1::5])
    labels, inertia = ld.assign_nearest(X, Y, "euclidean")
    monkeypatch.setenv("MSMB200_ASSIGN_EXACT", "1")
    ref_labels, ref_inertia = ld.assign_nearest(X, Y, "euclidean")
    np.testing.assert_array_equal(labels, ref_labels)
    assert abs(inertia - ref_inertia) <= 1e-12 * abs(ref_inertia)
    monkeypatch.delenv("MSMB200_ASSIGN_EXACT")
    monkeypatch.setenv("MSMB200_ASSIGN_SIMT", "1")
    simt_labels, _ = ld.assign_nearest(X, Y, "sqeuclidean")
    np.testing.assert_array_equal(simt_labels, ref_labels)


@pytest.mark.parametrize("n,d,k", [(6000, 128, 300), (5000, 256, 100), (8000, 64, 600), (9000, 16, 2000),
                                   (5000, 100, 129), (148 * 128 * 2 + 77, 128, 256), (4500, 256, 2)])
def test_streamed_centres_filter_labels_equal_exact_engine(n, d, k, monkeypatch):
    # K3 with streamed centre chunks (assign_umma_stream_kernel: the centre table does not fit shared
    # memory -- every d > 64, k = 2000 at d = 16): several chunks per frame tile, more frame tiles than
    # SMs, single- and double-buffered frame operands; labels of the float64 scan, ties included
    from msmbuilder_b200 import libdistance as ld, _lib
    assert _lib.load().msmb200_assign_engine(n, k, d) == 1
    rs = np.random.RandomState(n + d + k)
    X = (rs.randn(n, d) * 10.0 ** rs.uniform(-2, 2, size=d)).astype(np.float32)
    Y = X[rs.choice(n, k, replace=False)].copy()
    Y[k // 2] = Y[0]                                   # duplicate centre: lowest index wins
    X[::7] = np.round(X[::7])                          # lattice points: exact ties between centres
    Y[1::5] = np.round(Y[1::5])
    labels, inertia = ld.assign_nearest(X, Y, "euclidean")
    monkeypatch.setenv("MSMB200_ASSIGN_EXACT", "1")
    ref_labels, ref_inertia = ld.assign_nearest(X, Y, "euclidean")
    np.testing.assert_array_equal(labels, ref_labels)
    assert abs(inertia - ref_inertia) <= 1e-12 * abs(ref_inertia)
    monkeypatch.delenv("MSMB200_ASSIGN_EXACT")
    # a resident-size problem forced through the streamed kernel gives the same labels too
    if d <= 64 and k <= 256:
        return
    sq_labels, _ = ld.assign_nearest(X, Y, "sqeuclidean")
    np.testing.assert_array_equal(sq_labels, ref_labels)


def test_streamed_kernel_on_a_resident_size_problem(monkeypatch):
    from msmbuilder_b200 import libdistance as ld
    rs = np.random.RandomState(3)
    X = rs.randn(20000, 16).astype(np.float32)
    Y = X[rs.choice(len(X), 500, replace=False)].copy()
    labels, _ = ld.assign_nearest(X, Y, "euclidean")
    monkeypatch.setenv("MSMB200_ASSIGN_STREAM", "1")
    s_labels, _ = ld.assign_nearest(X, Y, "euclidean")
    np.testing.assert_array_equal(s_labels, labels)
    ref_labels, _ = lo.assign_nearest(X[:3000], Y, "euclidean")
    np.testing.assert_array_equal(s_labels[:3000], ref_labels)


def test_streamed_centres_filter_vs_reference_cpp():
    from msmbuilder_b200 import libdistance as ld
    rs = np.random.RandomState(8)
    X = rs.randn(4096, 128).astype(np.float32)
    Y = rs.randn(40, 128).astype(np.float32)
    labels, inertia = ld.assign_nearest(X, Y, "euclidean")
    ref_labels, ref_inertia = lo.assign_nearest(X, Y, "euclidean")
    np.testing.assert_array_equal(labels, ref_labels)
    assert abs(inertia - ref_inertia) <= 1e-11 * abs(ref_inertia)


def test_tensor_core_filter_vs_reference_cpp():
    # against the compiled reference itself (oracle/_ref) on a size it finishes quickly
    from msmbuilder_b200 import libdistance as ld
    rs = np.random.RandomState(5)
    X = rs.randn(8192, 16).astype(np.float32)
    Y = rs.randn(64, 16).astype(np.float32)
    labels, inertia = ld.assign_nearest(X, Y, "euclidean")
    ref_labels, ref_inertia = lo.assign_nearest(X, Y, "euclidean")
    np.testing.assert_array_equal(labels, ref_labels)
    assert abs(inertia - ref_inertia) <= 1e-11 * abs(ref_inertia)
