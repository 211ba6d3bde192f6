"""GPU parity of the look-ahead k-centers (csrc/kcenters_lookahead.cu): centre ids, labels and
distances must equal the pass-per-centre path (msmb200_kcenters_pass, itself pinned to the
reference loop kcenters.py:79-102 by test_gpu_cluster.py) bit for bit, on data where the chain
certifies many centres per pass and on data where it cannot."""
import numpy as np
import pytest

from oracle import cluster_oracle as co

pytestmark = pytest.mark.gpu


def _data(kind, n, d, seed):
    rs = np.random.RandomState(seed)
    if kind == "gauss":
        X = rs.randn(n, d)
    elif kind == "clustered":         # tight blobs: every new centre drags a whole blob down
        cen = rs.randn(6, d) * 10
        X = cen[rs.randint(0, 6, n)] + 0.05 * rs.randn(n, d)
    elif kind == "integers":          # exact ties everywhere
        X = rs.randint(-2, 3, size=(n, d)).astype(np.float64)
    elif kind == "duplicates":        # repeated frames (zero distances, equal maxima)
        base = rs.randn(max(n // 7, 1), d)
        X = base[rs.randint(0, len(base), n)]
    elif kind == "offset":            # large common offset: float32 filter margins
        X = 1000.0 + rs.randn(n, d)
    else:
        raise ValueError(kind)
    return np.ascontiguousarray(X, dtype=np.float32)


def _both(X, k, metric="euclidean", seed_index=None):
    import torch
    from msmbuilder_b200 import _kernels as K
    Xd = torch.from_numpy(X).cuda()
    seed_index = (len(X) // 3) if seed_index is None else seed_index
    assert K.lookahead_supported(Xd, metric)
    sa, sb = {}, {}
    a = K.kcenters_fit(Xd, k, metric, seed_index, lookahead=True, stats=sa)
    b = K.kcenters_fit(Xd, k, metric, seed_index, lookahead=False, stats=sb)
    ids_a, ids_b = a[0].cpu().numpy(), b[0].cpu().numpy()
    np.testing.assert_array_equal(ids_a, ids_b)
    np.testing.assert_array_equal(a[2].cpu().numpy(), b[2].cpu().numpy())       # labels
    np.testing.assert_array_equal(a[1].cpu().numpy(), b[1].cpu().numpy())       # distances, bit for bit
    assert sb["passes"] == k and sa["passes"] <= k
    return sa["passes"], ids_a


@pytest.mark.parametrize("kind", ["gauss", "clustered", "integers", "duplicates", "offset"])
@pytest.mark.parametrize("n,d,k", [(50000, 256, 8), (30011, 64, 25), (20000, 16, 40), (4099, 128, 12),
                                   (9000, 512, 6), (7001, 32, 17)])
def test_lookahead_equals_pass_per_centre(kind, n, d, k):
    _both(_data(kind, n, d, seed=n + d + k), k)


@pytest.mark.parametrize("n,k", [(1, 1), (3, 3), (5, 3), (40, 10), (40, 40), (257, 8), (1023, 30)])
def test_lookahead_tiny_inputs(n, k):
    _both(_data("gauss", n, 32, seed=n), k, seed_index=n - 1)


def test_lookahead_sqeuclidean_and_oracle():
    X = _data("gauss", 20000, 16, seed=3)
    _both(X, 25, metric="sqeuclidean")
    _, ids = _both(X, 25, seed_index=int(np.random.RandomState(4).randint(0, len(X))))
    r = co.kcenters_fit(X, 25, "euclidean", random_state=4)
    assert list(ids) == r["cluster_ids_"]


def test_lookahead_needs_few_passes_on_the_bench_like_data():
    # 256-d near-isotropic frames: the far points are far from each other, so one chain
    # certifies all remaining centres: 2 reads of the frames for k = 8 instead of 8
    import torch
    from msmbuilder_b200.synthetic import ar1_device
    from msmbuilder_b200 import _kernels as K
    X = ar1_device(4, 50000, 256, seed=1000)
    sa = {}
    a = K.kcenters_fit(X, 8, "euclidean", 12345, lookahead=True, stats=sa)
    b = K.kcenters_fit(X, 8, "euclidean", 12345, lookahead=False)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and torch.equal(a[2], b[2])
    assert sa["passes"] <= 3


def test_estimator_takes_the_lookahead_path():
    from msmbuilder_b200.cluster import KCenters
    X = _data("gauss", 30000, 64, seed=9)
    m = KCenters(n_clusters=12, random_state=1).fit([X[:10000], X[10000:]])
    r = co.kcenters_fit(X, 12, "euclidean", random_state=1)
    assert m.cluster_ids_ == r["cluster_ids_"]
    np.testing.assert_array_equal(np.concatenate(m.labels_), r["labels_"])
    np.testing.assert_allclose(np.concatenate(m.distances_), r["distances_"], rtol=1e-13)
    np.testing.assert_array_equal(m.cluster_centers_, X[r["cluster_ids_"]])
