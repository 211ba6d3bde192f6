"""GPU: metric='rmsd' (QCP kernels) vs the float64 oracle.  PARITY UNPINNED against
the reference (mdtraj's libtheobald is absent): tolerance 1e-5 is what the
reference's own RMSD tests accept (msmbuilder/tests/test_libdistance.py:155,163)."""
import numpy as np
import pytest

from oracle import rmsd_oracle as ro
from oracle import cluster_oracle as co
from msmbuilder_b200.synthetic import rmsd_conformations_numpy

pytestmark = pytest.mark.gpu
TOL = 1e-5


def assert_rmsd_close(got, ref, G_scale, n_atoms):
    """RMSD comes out of msd = (G_a + G_b - 2 lambda) / n with float32 traces (as in
    the reference, libdistance.pyx:336-341), so the ERROR lives in msd: about
    6e-8 * (G_a + G_b) / n.  Near-zero distances therefore carry an absolute error
    of ~sqrt(that) (1e-4) in both implementations; compare squared distances."""
    tol_msd = 4e-7 * G_scale / n_atoms
    np.testing.assert_allclose(np.asarray(got) ** 2, np.asarray(ref) ** 2, rtol=2e-5, atol=tol_msd)


def test_center_dist_pdist_assign():
    from msmbuilder_b200 import libdistance as ld
    xyz, _ = rmsd_conformations_numpy(300, n_atoms=37, n_templates=6, seed=0)
    c, G = ro.center_and_trace(xyz)
    D = ro.rmsd_qcp(c, c[:9], G, G[:9])
    Gs = 2 * float(G.max())
    assert_rmsd_close(ld.cdist(xyz, xyz[:9], "rmsd"), D, Gs, 37)
    assert_rmsd_close(ld.dist(xyz, xyz[4], "rmsd"), D[:, 4], Gs, 37)
    idx = np.array([5, 1, 200, 33, 5])
    assert_rmsd_close(ld.pdist(xyz, "rmsd", idx), ro.pdist(c, idx, G), Gs, 37)
    labels, inertia = ld.assign_nearest(xyz, xyz[:9], "rmsd")
    Ds = np.sort(D, axis=1)
    clear = (Ds[:, 1] - Ds[:, 0]) > 10 * TOL
    np.testing.assert_array_equal(labels[clear], D.argmin(1)[clear])
    assert abs(inertia - D.min(1).sum()) < 300 * 2e-4


def test_kcenters_rmsd_vs_oracle():
    from msmbuilder_b200.cluster import KCenters
    xyz, which = rmsd_conformations_numpy(2000, n_atoms=20, n_templates=8, seed=1, noise=0.02)
    c, G = ro.center_and_trace(xyz)
    r = co.kcenters_fit(c, 8, "rmsd", random_state=0,
                        dist_fn=lambda X_, y_: ro.rmsd_qcp(X_, y_[None])[:, 0])
    m = KCenters(n_clusters=8, metric="rmsd", random_state=0).fit([xyz[:1200], xyz[1200:]])
    # 8 well separated templates: the 8 centres must be one frame of each template
    assert sorted(which[m.cluster_ids_]) == list(range(8))
    assert m.cluster_ids_ == r["cluster_ids_"]
    np.testing.assert_array_equal(np.concatenate(m.labels_), r["labels_"])
    assert_rmsd_close(np.concatenate(m.distances_), r["distances_"], 2 * float(G.max()), 20)
    assert m.cluster_centers_.shape == (8, 20, 3)
    np.testing.assert_array_equal(np.concatenate(m.predict([xyz[:1200], xyz[1200:]])), r["labels_"])


def test_minibatch_kmedoids_rmsd_runs():
    from msmbuilder_b200.cluster import MiniBatchKMedoids
    xyz, which = rmsd_conformations_numpy(600, n_atoms=15, n_templates=4, seed=2, noise=0.01)
    m = MiniBatchKMedoids(n_clusters=4, batch_size=50, metric="rmsd", random_state=0).fit([xyz])
    lab = m.labels_[0]
    # frames of one template share a label
    for t in range(4):
        assert len(set(lab[which == t])) == 1
