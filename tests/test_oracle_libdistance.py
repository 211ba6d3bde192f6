"""CPU: pin the libdistance / k-medoids oracle.

(1) bit-for-bit against the reference's own C++ (oracle/_ref/libref.so);
(2) against scipy.spatial.distance the way the reference's tests do
    (msmbuilder/tests/test_libdistance.py:28-73,115-148,176-196,231-275);
(3) the k-medoids known-answer tests of msmbuilder/tests/test_kmedoids.py.
"""
import numpy as np
import pytest
import scipy.spatial.distance

from oracle import libdistance_oracle as lo

METRICS = list(lo.VECTOR_METRICS)
needs_ref = pytest.mark.skipif(not lo.have_reference(), reason="oracle/_ref not built")


def _data(seed, n, d, dtype, metric):
    rs = np.random.RandomState(seed)
    X = rs.randn(n, d)
    if metric in ("hamming", "jaccard"):
        X = np.round(X)
    X[rs.rand(n, d) > 0.9] = 0.0
    return X.astype(dtype)


@needs_ref
@pytest.mark.parametrize("metric", METRICS)
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_port_equals_reference_bitwise(metric, dtype):
    X = _data(0, 60, 7, dtype, metric)
    Y = _data(1, 9, 7, dtype, metric)
    X[3] = 0
    Y[2] = 0
    idx = np.random.RandomState(2).randint(0, 60, 17)
    for f, args in [(lo.cdist, (X, Y, metric)), (lo.pdist, (X, metric)),
                    (lo.pdist, (X, metric, idx)), (lo.dist, (X, Y[1], metric)),
                    (lo.dist, (X, Y[1], metric, idx))]:
        a = f(*args, impl="port")
        b = f(*args, impl="reference")
        assert np.array_equal(a, b, equal_nan=True), (f.__name__, metric)
    for rows in (None, idx):
        a = lo.assign_nearest(X, Y, metric, rows, impl="port")
        b = lo.assign_nearest(X, Y, metric, rows, impl="reference")
        assert np.array_equal(a[0], b[0])
        assert a[1] == b[1] or (np.isnan(a[1]) and np.isnan(b[1]))
    pairs = np.random.RandomState(3).randint(0, 60, (25, 2))
    sa, sb = lo.sumdist(X, metric, pairs, impl="port"), lo.sumdist(X, metric, pairs, impl="reference")
    assert sa == sb or (np.isnan(sa) and np.isnan(sb))


@pytest.mark.parametrize("metric", METRICS)
@pytest.mark.parametrize("dtype,places", [(np.float32, 5), (np.float64, 10)])
def test_matches_scipy(metric, dtype, places):
    # reference: test_libdistance.py:115-148,176-196 (decimal=5 for f32, 10 for f64)
    X = _data(4, 40, 6, dtype, metric) + (0 if metric in ("hamming", "jaccard") else 0.1)
    Y = _data(5, 8, 6, dtype, metric) + (0 if metric in ("hamming", "jaccard") else 0.1)
    X64, Y64 = X.astype(np.float64), Y.astype(np.float64)
    if metric == "jaccard":
        # SciPy >= 1.2 booleanises its inputs; the reference keeps the numeric
        # definition of its era (distance_kernels.h:196-222): mismatches among
        # positions where either vector is non-zero.
        def jac(u, v):
            nz = (u != 0) | (v != 0)
            return ((u != v) & nz).sum() / nz.sum()
        ref = np.array([[jac(u, v) for v in Y64] for u in X64])
        refp = np.array([jac(X64[i], X64[j]) for i in range(len(X64)) for j in range(i + 1, len(X64))])
    else:
        ref = scipy.spatial.distance.cdist(X64, Y64, metric=metric)
        refp = scipy.spatial.distance.pdist(X64, metric=metric)
    got = lo.cdist(X, Y, metric)
    np.testing.assert_array_almost_equal(got, ref, decimal=places)
    np.testing.assert_array_almost_equal(lo.pdist(X, metric), refp, decimal=places)


@pytest.mark.parametrize("metric", METRICS)
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_assign_equals_cdist_argmin(metric, dtype):
    # reference: test_libdistance.py:28-73
    X = _data(6, 50, 5, dtype, metric)
    Y = _data(7, 7, 5, dtype, metric)
    idx = np.random.RandomState(8).randint(0, 50, 11)
    for rows in (None, idx):
        labels, inertia = lo.assign_nearest(X, Y, metric, rows)
        D = lo.cdist(X if rows is None else X[rows], Y, metric)
        Dn = np.where(np.isnan(D), np.inf, D)
        np.testing.assert_array_equal(labels[~np.isnan(D).all(1)], Dn.argmin(1)[~np.isnan(D).all(1)])
        if np.isfinite(inertia):
            np.testing.assert_almost_equal(inertia, D[np.arange(len(D)), labels].sum(), decimal=8)


def test_unknown_metric_and_dtype_errors():
    X = np.zeros((3, 2), dtype=np.float32)
    with pytest.raises(ValueError):
        lo.cdist(X, X, "nope")
    with pytest.raises(TypeError):
        lo.cdist(X, X.astype(np.float64), "euclidean")


# ---- k-medoids (reference: msmbuilder/tests/test_kmedoids.py) ----------------------
def test_condensed_index():
    # test_kmedoids.py 'test_index': ix(i, j, n) addresses scipy's squareform order
    n = 9
    D = np.random.RandomState(0).rand(n, n)
    D = D + D.T
    np.fill_diagonal(D, 0)
    dm = scipy.spatial.distance.squareform(D)
    for i in range(n):
        for j in range(n):
            if i != j:
                assert dm[lo.condensed_index(i, j, n)] == D[i, j]


def test_contigify_ids():
    # test_kmedoids.py:15-29
    ids, m = lo.contigify_ids(np.array([4, 4, 9, 2, 9, 4]))
    np.testing.assert_array_equal(ids, [0, 0, 1, 2, 1, 0])
    assert m == {4: 0, 9: 1, 2: 2}
    ids, m = lo.contigify_ids(np.array([0, 1, 2]))
    np.testing.assert_array_equal(ids, [0, 1, 2])


def test_kmedoids_obvious_two_blobs():
    # test_kmedoids.py 'test_obvious_clustering': two well separated blobs are recovered
    rs = np.random.RandomState(1)
    X = np.concatenate([rs.randn(20, 2) * 0.1, rs.randn(20, 2) * 0.1 + 10.0])
    dm = lo.pdist(X, "euclidean")
    init = rs.randint(0, 2, 40)
    init[0], init[20] = 0, 1
    ids, err, found = lo.kmedoids(2, dm, 0, init)
    assert found == 1
    assert len(set(ids[:20])) == 1 and len(set(ids[20:])) == 1 and ids[0] != ids[39]
    # inertia definition (test_kmedoids.py 'test_inertia'): sum of distances to the medoid
    D = scipy.spatial.distance.squareform(dm)
    np.testing.assert_almost_equal(err, sum(D[i, ids[i]] for i in range(40)))


@needs_ref
def test_kmedoids_port_equals_reference():
    rs = np.random.RandomState(2)
    for trial in range(25):
        n, k = rs.randint(8, 70), rs.randint(2, 7)
        X = rs.randn(n, 3)
        if trial % 5 == 0:
            X = np.round(X)           # exact ties
        dm = lo.pdist(X, "euclidean")
        cid = rs.randint(0, k, n)
        cid[:k] = np.arange(k)
        a = lo.kmedoids(k, dm, 0, cid, impl="port")
        b = lo.kmedoids(k, dm, 0, cid, impl="reference")
        np.testing.assert_array_equal(a[0], b[0])
        assert a[1] == b[1] and a[2] == b[2]
        ca, cb = lo.contigify_ids(a[0].copy(), impl="port"), lo.contigify_ids(b[0].copy(), impl="reference")
        np.testing.assert_array_equal(ca[0], cb[0])
        assert ca[1] == cb[1]
