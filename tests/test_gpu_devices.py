"""GPU (>= 2 devices): `devices=` -- one process, one thread per GPU -- gives what one GPU gives."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _need_two():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")


def test_tica_devices_matches_one_gpu():
    _need_two()
    from msmbuilder_b200.decomposition import tICA
    from msmbuilder_b200.synthetic import ar1_numpy
    seqs = [s[:n] for s, n in zip(ar1_numpy(7, 30000, 64, seed=3), [30000, 1200, 25000, 7, 18000, 30000, 999])]
    a = tICA(n_components=4, lag_time=10, engine="simt_f64").fit(seqs)
    b = tICA(n_components=4, lag_time=10, engine="simt_f64", devices="all").fit(seqs)
    assert a.n_observations_ == b.n_observations_ and a.n_sequences_ == b.n_sequences_
    np.testing.assert_allclose(b._outer_0_to_T_lagged, a._outer_0_to_T_lagged, rtol=1e-12, atol=1e-9)
    np.testing.assert_allclose(b.eigenvalues_, a.eigenvalues_, rtol=0, atol=1e-12)
    c = tICA(n_components=4, lag_time=10, devices=[0, 1]).fit(seqs)       # default engine
    np.testing.assert_allclose(c.eigenvalues_, a.eigenvalues_, rtol=0, atol=1e-5)


@pytest.mark.parametrize("metric,dtype", [("euclidean", np.float32), ("cityblock", np.float64)])
def test_kcenters_devices_matches_one_gpu(metric, dtype):
    _need_two()
    from msmbuilder_b200.cluster import KCenters
    rs = np.random.RandomState(1)
    seqs = [rs.randn(n, 32).astype(dtype) for n in (20000, 5000, 33333, 12000)]
    a = KCenters(n_clusters=12, metric=metric, random_state=4).fit(seqs)
    b = KCenters(n_clusters=12, metric=metric, random_state=4, devices="all").fit(seqs)
    assert a.cluster_ids_ == b.cluster_ids_
    np.testing.assert_array_equal(a.cluster_centers_, b.cluster_centers_)
    for x, y in zip(a.labels_, b.labels_):
        np.testing.assert_array_equal(x, y)
    for x, y in zip(a.distances_, b.distances_):
        np.testing.assert_array_equal(x, y)
    assert abs(a.inertia_ - b.inertia_) <= 1e-9 * abs(a.inertia_)


def test_kcenters_rmsd_devices_matches_one_gpu():
    _need_two()
    from msmbuilder_b200.cluster import KCenters
    from msmbuilder_b200.synthetic import rmsd_conformations_numpy
    xyz, _ = rmsd_conformations_numpy(30000, n_atoms=20, n_templates=10, seed=2, noise=0.03)
    a = KCenters(n_clusters=25, metric="rmsd", random_state=0).fit([xyz[:11000], xyz[11000:]])
    b = KCenters(n_clusters=25, metric="rmsd", random_state=0, devices=[0, 1]).fit([xyz[:11000], xyz[11000:]])
    assert a.cluster_ids_ == b.cluster_ids_
    for x, y in zip(a.labels_, b.labels_):
        np.testing.assert_array_equal(x, y)
    for x, y in zip(a.distances_, b.distances_):
        np.testing.assert_array_equal(x, y)
