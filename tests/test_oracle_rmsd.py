"""CPU: the RMSD oracle is PARITY-UNPINNED against the reference (mdtraj absent;
see oracle/rmsd_oracle.py).  What can be checked: the two independent float64
routes (QCP eigen-solve vs Kabsch SVD) agree, and the metric behaves like an
RMSD (rotation/translation invariance, identity, symmetry)."""
import numpy as np

from oracle import rmsd_oracle as ro
from msmbuilder_b200.synthetic import rmsd_conformations_numpy


def test_qcp_equals_kabsch():
    xyz, _ = rmsd_conformations_numpy(40, n_atoms=17, n_templates=4, seed=0)
    c, G = ro.center_and_trace(xyz)
    a = ro.rmsd_qcp(c[:25], c[25:])
    b = ro.rmsd_kabsch(c[:25], c[25:])
    np.testing.assert_allclose(a, b, atol=2e-6)


def test_invariances():
    xyz, which = rmsd_conformations_numpy(30, n_atoms=12, n_templates=3, seed=1, noise=0.0)
    c, G = ro.center_and_trace(xyz)
    D = ro.rmsd_qcp(c, c, G, G)
    same = which[:, None] == which[None, :]
    assert D[same].max() < 2e-3          # same template, only rotated / translated (float32 coords)
    assert D[~same].min() > 0.05
    np.testing.assert_allclose(D, D.T, atol=1e-6)


def test_assign_and_pdist_shapes():
    xyz, _ = rmsd_conformations_numpy(20, n_atoms=10, n_templates=3, seed=2)
    c, G = ro.center_and_trace(xyz)
    labels, inertia = ro.assign_nearest(c, c[:4], G, G[:4])
    assert labels.shape == (20,) and (labels[:4] == np.arange(4)).all() and inertia >= 0
    p = ro.pdist(c, X_indices=[0, 3, 5, 7], GX=G)
    assert p.shape == (6,)
