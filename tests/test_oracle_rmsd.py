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


def test_published_qcp_known_answer():
    # Theobald's qcprot main.c fragments: "QCP rmsd: 0.719106" -- both float64 routes, the Newton
    # iteration on a float64 M and the float32 restatement reproduce the published digits
    a, ga = ro.center_and_trace(ro.QCPROT_FRAG_A[None])
    b, gb = ro.center_and_trace(ro.QCPROT_FRAG_B[None])
    assert abs(ro.rmsd_qcp(a, b, ga, gb)[0, 0] - ro.QCPROT_RMSD) < 2e-6
    assert abs(ro.rmsd_kabsch(a, b)[0, 0] - ro.QCPROT_RMSD) < 2e-6
    M = a[0].astype(np.float64).T @ b[0].astype(np.float64)
    assert abs(np.sqrt(ro.msd_from_M_and_G(M, ga[0], gb[0], 7)) - ro.QCPROT_RMSD) < 2e-6
    assert abs(float(ro.rmsd_theobald_f32(a, b, ga, gb)[0, 0]) - ro.QCPROT_RMSD) < 5e-6
    # and the optimal rotation is the published one (Kabsch route)
    ac = ro.QCPROT_FRAG_A - ro.QCPROT_FRAG_A.mean(0)
    bc = ro.QCPROT_FRAG_B - ro.QCPROT_FRAG_B.mean(0)
    U, S, Vt = np.linalg.svd(ac.T @ bc)
    R = U @ np.diag([1.0, 1.0, np.sign(np.linalg.det(U @ Vt))]) @ Vt
    np.testing.assert_allclose(R, ro.QCPROT_ROTATION, atol=2e-8)


def test_float32_restatement_envelope():
    # the float32 accumulation the reference calls differs from the float64 routes by the float32
    # rounding of M only: <= 1e-5 on the msd scale, which is what the reference's tests accept
    # (tests/test_libdistance.py:155,163 decimal=5)
    xyz, _ = rmsd_conformations_numpy(24, n_atoms=100, n_templates=4, seed=3)
    c, G = ro.center_and_trace(xyz)
    d64 = ro.rmsd_qcp(c[:16], c[16:], G[:16], G[16:])
    d32 = ro.rmsd_theobald_f32(c[:16], c[16:], G[:16], G[16:]).astype(np.float64)
    scale = 2.0 * float(G.max()) / 100
    assert np.abs(d32 ** 2 - d64 ** 2).max() < 1e-5 * scale
