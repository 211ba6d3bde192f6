"""Pins oracle/msm_oracle.py (transition counting) against
  * the reference's own `_transition_counts`, loaded verbatim (build container only),
  * the known answers of the reference's tests/test_transition_counts.py,
  * tests/golden/msm_counts.npz.
"""
import os
import warnings

import numpy as np
import pytest

from oracle import msm_oracle as mo
from oracle import ref_loader

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "msm_counts.npz")


def random_label_sequences(seed, n_seq=7, n_states=13, gaps=False):
    rs = np.random.RandomState(seed)
    out = []
    for i in range(n_seq):
        n = int(rs.randint(1, 400))
        y = rs.randint(0, n_states, size=n)
        if gaps:
            y = y * 3 + 100            # non-contiguous labels
        out.append(y.astype(np.int64))
    return out


def test_known_answers_of_the_reference_tests():
    # tests/test_transition_counts.py:13-26,38-46,66-78 and the docstring core.py:517-532
    c, m = mo.transition_counts([np.arange(10)])
    np.testing.assert_array_equal(c, np.eye(10, k=1))
    assert list(m.keys()) == list(range(10)) and list(m.values()) == list(range(10))
    c, m = mo.transition_counts([range(10)], lag_time=2)
    np.testing.assert_array_equal(c, 0.5 * np.eye(10, k=2))
    c, m = mo.transition_counts([[100000000, 100000000, 100000001, 100000001]])
    np.testing.assert_array_equal(c, np.array([[1., 1.], [0., 1.]]))
    assert m == {100000000: 0, 100000001: 1}
    c, m = mo.transition_counts([[0, 0, 0, 1, 1]])
    np.testing.assert_array_equal(c, np.array([[2., 1.], [0., 1.]]))
    c, m = mo.transition_counts([[100, 200, 300]])
    np.testing.assert_array_equal(c, np.eye(3, k=1))
    C, _ = mo.transition_counts([np.arange(6)], lag_time=3)
    np.testing.assert_array_almost_equal(C, np.eye(6, k=3) / 3)
    X = np.arange(10)
    C1, m1 = mo.transition_counts([X], lag_time=3, sliding_window=False)
    C2, m2 = mo.transition_counts([X[::3]], sliding_window=True)
    np.testing.assert_array_almost_equal(C1, C2)
    assert m1 == m2


def test_nan_and_short():
    # tests/test_transition_counts.py:48-60
    c, m = mo.transition_counts([[0]])
    assert c.shape == (1, 1) and c[0, 0] == 0
    c, m = mo.transition_counts([[0, np.nan]])
    assert m == {0: 0}
    np.testing.assert_array_equal(c, np.zeros((1, 1)))
    c, m = mo.transition_counts([[np.nan]])
    assert m == {}
    np.testing.assert_array_equal(c, np.zeros((0, 0)))


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present")
@pytest.mark.parametrize("lag,sliding,gaps", [(1, True, False), (3, True, False), (4, False, False),
                                              (2, True, True), (5, False, True)])
def test_against_reference_verbatim(lag, sliding, gaps):
    ref = ref_loader.load_transition_counts()
    seqs = random_label_sequences(lag * 7 + gaps, gaps=gaps)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        c_ref, m_ref = ref(seqs, lag_time=lag, sliding_window=sliding)
    c, m = mo.transition_counts(seqs, lag_time=lag, sliding_window=sliding)
    np.testing.assert_array_equal(c, c_ref)
    assert {int(k): v for k, v in m_ref.items()} == m


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present")
def test_against_reference_with_nan():
    ref = ref_loader.load_transition_counts()
    rs = np.random.RandomState(3)
    seqs = []
    for n in (50, 1, 333):
        y = rs.randint(0, 6, size=n).astype(np.float64)
        y[rs.rand(n) < 0.1] = np.nan
        seqs.append(y)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        c_ref, m_ref = ref(seqs, lag_time=2)
    c, m = mo.transition_counts(seqs, lag_time=2)
    np.testing.assert_array_equal(c, c_ref)
    assert {float(k): v for k, v in m_ref.items()} == m


def test_golden():
    g = np.load(GOLDEN)
    for case in range(int(g["n_cases"])):
        lens = g["lens_%d" % case]
        flat = g["labels_%d" % case]
        seqs = np.split(flat, np.cumsum(lens)[:-1])
        c, m = mo.transition_counts(seqs, lag_time=int(g["lag_%d" % case]),
                                    sliding_window=bool(g["sliding_%d" % case]))
        np.testing.assert_array_equal(c, g["counts_%d" % case])
        np.testing.assert_array_equal(np.array(sorted(m.keys())), g["classes_%d" % case])
