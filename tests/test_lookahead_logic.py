"""CPU: the look-ahead k-centers LOGIC (lane top-two -> candidates + tau -> certified chain -> fused
update) reproduces the reference loop (kcenters.py:79-102 via the oracle) for every lane count,
candidate cap and chain cap -- including caps so small that nothing can be certified, exact ties,
duplicates and k > number of distinct frames.  Runs the same driver as the GPU path
(_kernels.kcenters_fit_lookahead) on the host stand-in of tests/_lookahead_host.py."""
import numpy as np
import pytest

from _lookahead_host import HostLookahead
from msmbuilder_b200 import _kernels as K
from oracle import cluster_oracle as co


def _ref(X, k, seed):
    class FixedSeed(object):
        def randint(self, lo_, hi):
            return seed
    orig = co.check_random_state
    co.check_random_state = lambda rs_: FixedSeed()
    try:
        return co.kcenters_fit(X, k, "euclidean")
    finally:
        co.check_random_state = orig


def _data(kind, n, d, seed):
    rs = np.random.RandomState(seed)
    if kind == "gauss":
        X = rs.randn(n, d)
    elif kind == "clustered":
        X = (rs.randn(4, d) * 6)[rs.randint(0, 4, n)] + 0.05 * rs.randn(n, d)
    elif kind == "integers":
        X = rs.randint(-1, 2, size=(n, d)).astype(np.float64)
    else:   # duplicates
        base = rs.randn(max(n // 5, 1), d)
        X = base[rs.randint(0, len(base), n)]
    return np.ascontiguousarray(X, dtype=np.float32)


@pytest.mark.parametrize("kind", ["gauss", "clustered", "integers", "duplicates"])
@pytest.mark.parametrize("n_lanes,t_cap,j_cap", [(1, 2, 1), (3, 2, 4), (7, 5, 4), (16, 8, 16), (64, 64, 3), (200, 512, 16)])
def test_lookahead_logic_equals_reference(kind, n_lanes, t_cap, j_cap):
    for n, d, k in ((60, 3, 9), (257, 5, 20), (31, 2, 31)):
        X = _data(kind, n, d, seed=n + n_lanes)
        seed = n // 2
        ref = _ref(X, k, seed)
        st = HostLookahead(X, 0, n_lanes=n_lanes, t_cap=t_cap, j_cap=j_cap)
        stats = {}
        ids, rows, distances, labels = K.kcenters_fit_lookahead(None, k, "euclidean", seed, stats=stats, state=st)
        assert list(ids.numpy()) == ref["cluster_ids_"], (kind, n, n_lanes, t_cap, j_cap)
        np.testing.assert_array_equal(labels.numpy(), ref["labels_"])
        np.testing.assert_array_equal(distances.numpy(), ref["distances_"])
        np.testing.assert_array_equal(rows.numpy(), X[ref["cluster_ids_"]])
        assert 1 <= stats["passes"] <= k


def test_lookahead_saves_passes_when_it_can():
    # many lanes, generous caps, spread-out data: far fewer reads than centres
    X = _data("gauss", 4000, 24, seed=1)
    st = HostLookahead(X, 0, n_lanes=500, t_cap=256, j_cap=16)
    stats = {}
    ids, _, _, _ = K.kcenters_fit_lookahead(None, 12, "euclidean", 7, stats=stats, state=st)
    assert list(ids.numpy()) == _ref(X, 12, 7)["cluster_ids_"]
    assert stats["passes"] < 12
