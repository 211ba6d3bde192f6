"""CPU: pin the tICA oracle (oracle/tica_oracle.py).

(1) against the reference's own tica.py loaded verbatim (build container only);
(2) against tests/golden/tica_*.npz written from that verbatim run;
(3) the identity tests of msmbuilder/tests/test_decomposition.py:28-125 and
    msmbuilder/tests/test_utils.py:59-79 that need no external data.
"""
import glob
import os
import warnings

import numpy as np
import pytest

from oracle import ref_loader
from oracle.tica_oracle import TicaOracle
from msmbuilder_b200.synthetic import ar1_numpy

needs_ref = pytest.mark.skipif(not ref_loader.available(), reason="/root/reference absent")


def _golden_inputs(g):
    seqs = ar1_numpy(int(g["n_seq"]), int(g["length"]), int(g["D"]), seed=int(g["seed"]),
                     dtype=np.dtype(str(g["dtype"])))
    for n_short in g["short"]:
        seqs.insert(1, seqs[0][:int(n_short)].copy())
    return seqs


def _golden_files(golden_dir):
    return sorted(glob.glob(os.path.join(golden_dir, "tica_*.npz")))


def test_golden_files_present(golden_dir):
    assert len(_golden_files(golden_dir)) >= 4


@pytest.mark.parametrize("name", ["tica_d6_lag3", "tica_d16_lag10", "tica_d64_lag10_f64",
                                  "tica_d256_lag10"])
def test_oracle_matches_golden(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    seqs = _golden_inputs(g)
    shrink = None if np.isnan(g["shrinkage"]) else float(g["shrinkage"])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = TicaOracle(n_components=int(g["k"]), lag_time=int(g["lag"]), shrinkage=shrink).fit(seqs)
    assert m.n_observations_ == int(g["n_observations"])
    assert m.n_sequences_ == int(g["n_sequences"])
    # same NumPy calls in the same order => identical to the last bit here
    np.testing.assert_array_equal(m._outer_0_to_T_lagged, g["C_tau"])
    np.testing.assert_array_equal(m._outer_0_to_TminusTau, g["C_00"])
    np.testing.assert_array_equal(m._outer_offset_to_T, g["C_tt"])
    np.testing.assert_array_equal(m._sum_0_to_T, g["S"])
    np.testing.assert_allclose(m.eigenvalues_, g["eigenvalues"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(m.means_, g["means"], rtol=0, atol=1e-14)
    proj = m.transform([seqs[0][:50]])[0]
    sign = np.sign((proj * g["proj50"]).sum(0))
    np.testing.assert_allclose(proj * sign, g["proj50"], atol=1e-8)


@needs_ref
def test_oracle_matches_verbatim_reference():
    tICA = ref_loader.load_tica()
    rs = np.random.RandomState(0)
    seqs = [rs.randn(300, 5).cumsum(0) * 0.05 + rs.randn(300, 5) for _ in range(3)] + [rs.randn(2, 5)]
    for shrink in (None, 0.0, 0.2):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            r = tICA(n_components=3, lag_time=4, shrinkage=shrink, kinetic_mapping=True).fit(seqs)
            o = TicaOracle(n_components=3, lag_time=4, shrinkage=shrink, kinetic_mapping=True).fit(seqs)
        np.testing.assert_array_equal(r._outer_0_to_T_lagged, o._outer_0_to_T_lagged)
        np.testing.assert_allclose(r.eigenvalues_, o.eigenvalues_, atol=1e-13)
        np.testing.assert_allclose(np.abs(r.transform(seqs[:1])[0]), np.abs(o.transform(seqs[:1])[0]),
                                   atol=1e-10)
        np.testing.assert_allclose(r.score(seqs[:2]), o.score(seqs[:2]), atol=1e-10)


# ---- reference identities (test_decomposition.py) -------------------------------------
def test_shapes():
    # test_decomposition.py:52-56
    m = TicaOracle(n_components=3).fit([np.random.RandomState(0).randn(10, 3)] + [np.random.RandomState(1).randn(10, 3)])
    assert m.eigenvalues_.shape == (3,) and m.eigenvectors_.shape == (3, 3) and m.components_.shape == (3, 3)


def test_singular():
    # test_decomposition.py:28-49: a repeated / all-zero column still solves (shrinkage rescues Sigma)
    rs = np.random.RandomState(0)
    X = rs.randn(100, 2)
    X = np.hstack([X, X[:, :1]])
    m = TicaOracle(n_components=2).fit([X])
    assert m.eigenvalues_.dtype == np.float64 and np.isfinite(m.eigenvalues_).all()
    Z = rs.randn(100, 3)
    Z[:, 0] = 0.0
    m = TicaOracle(n_components=2).fit([Z])
    assert np.isfinite(m.eigenvectors_).all()


def test_score_1():
    # test_decomposition.py:59-67: with shrinkage=0, score([X]) == eigenvalues.sum() == score_
    X = np.random.RandomState(0).randn(100, 5)
    for n in range(1, 5):
        m = TicaOracle(n_components=n, shrinkage=0).fit([X])
        np.testing.assert_almost_equal(m.eigenvalues_.sum(), m.score([X]))
        np.testing.assert_almost_equal(m.eigenvalues_.sum(), m.score_)


def test_multiple_components():
    # test_decomposition.py:80-98: raising n_components after fit re-solves lazily
    X = np.random.RandomState(0).randn(200, 6)
    m = TicaOracle(n_components=1).fit([X])
    t1 = m.transform([X])[0]
    m.n_components = 4
    t4 = m.transform([X])[0]
    assert t4.shape == (200, 4)
    np.testing.assert_allclose(np.abs(t1[:, 0]), np.abs(t4[:, 0]), atol=1e-10)


def test_kinetic_mapping():
    # test_decomposition.py:101-111
    X = np.random.RandomState(0).randn(200, 4).cumsum(0)
    a = TicaOracle(n_components=2, lag_time=2).fit([X])
    b = TicaOracle(n_components=2, lag_time=2, kinetic_mapping=True).fit([X])
    np.testing.assert_allclose(b.transform([X])[0], a.transform([X])[0] * a.eigenvalues_, atol=1e-10)


def test_subsampler_identity():
    # test_utils.py:59-79: tICA(lag=2) on X == tICA(lag=1) on X[::2] is NOT what is tested
    # there; the identity is tICA(lag=2).fit on [X] vs tICA(lag=1) on the two interleaved
    # sub-sequences: same lagged pairs, hence the same C_tau.
    X = np.random.RandomState(0).randn(401, 3).cumsum(0)
    a = TicaOracle(lag_time=2).fit([X])
    b = TicaOracle(lag_time=1).fit([X[0::2], X[1::2]])
    np.testing.assert_allclose(a._outer_0_to_T_lagged, b._outer_0_to_T_lagged, atol=1e-8)
    assert a.n_observations_ == b.n_observations_
