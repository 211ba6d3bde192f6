"""oracle/rmsd_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT.   *** PARITY UNPINNED ***

The RMSD metric's arithmetic is NOT in /root/reference: libdistance calls
mdtraj's libtheobald (`msd_atom_major`, `inplace_center_and_trace_atom_major`;
msmbuilder/libdistance/libdistance.pyx:67-72), a third-party dependency pinned
only as "mdtraj <=1.8" (devtools/conda-recipe/meta.yaml:21,29) and absent from
this image (no network).  The reference's own RMSD tests
(msmbuilder/tests/test_libdistance.py:76-112,151-173,199-228) need mdtraj plus
downloaded trajectories, so no golden vector of the reference can be produced
here: **parity for metric='rmsd' is unpinned** and is stated as such in
DESIGN.md.

What this file restates is the PUBLISHED algorithm that libtheobald implements
-- the quaternion characteristic polynomial (QCP) method of Theobald (Acta
Cryst. A61, 2005) and Liu, Agrafiotis & Theobald (J. Comput. Chem. 31, 2010) --
anchored on the reference's call sites:

  centre + trace   cluster/base.py:68,133-134 and libdistance.pyx:336-341:
                   every frame is translated to zero centroid and
                   G = sum_atoms |r|^2 is cached (float32) per frame;
  distance         libdistance.pyx:350-351,555-556:
                   rmsd = sqrtf(msd_atom_major(n, n, x_i, y_j, G_x[i], G_y[j], 0, NULL))
                   msd  = (G_a + G_b - 2*lambda_max) / n_atoms, lambda_max the
                   largest eigenvalue of the 4x4 key matrix built from M = A^T B.

Two independent float64 routes are provided and tested against each other:
`rmsd_qcp` (lambda_max by a dense symmetric eigen-solve of the key matrix) and
`rmsd_kabsch` (optimal rotation by SVD).  A tiny negative msd is clamped at 0
(the reference would produce NaN from sqrtf).
"""
import numpy as np


def center_and_trace(xyz):
    """(n_frames, n_atoms, 3) -> centred float32 copy and float32 traces G."""
    xyz = np.asarray(xyz)
    c = xyz.astype(np.float64)
    c = c - c.mean(axis=1, keepdims=True)
    traces = np.einsum("fad,fad->f", c, c)
    return c.astype(np.float32), traces.astype(np.float32)


def key_matrix(M):
    """4x4 symmetric key matrices for a batch of 3x3 inner-product matrices M = A^T B."""
    Sxx, Sxy, Sxz = M[..., 0, 0], M[..., 0, 1], M[..., 0, 2]
    Syx, Syy, Syz = M[..., 1, 0], M[..., 1, 1], M[..., 1, 2]
    Szx, Szy, Szz = M[..., 2, 0], M[..., 2, 1], M[..., 2, 2]
    K = np.empty(M.shape[:-2] + (4, 4), dtype=np.float64)
    K[..., 0, 0] = Sxx + Syy + Szz
    K[..., 0, 1] = K[..., 1, 0] = Syz - Szy
    K[..., 0, 2] = K[..., 2, 0] = Szx - Sxz
    K[..., 0, 3] = K[..., 3, 0] = Sxy - Syx
    K[..., 1, 1] = Sxx - Syy - Szz
    K[..., 1, 2] = K[..., 2, 1] = Sxy + Syx
    K[..., 1, 3] = K[..., 3, 1] = Szx + Sxz
    K[..., 2, 2] = -Sxx + Syy - Szz
    K[..., 2, 3] = K[..., 3, 2] = Syz + Szy
    K[..., 3, 3] = -Sxx - Syy + Szz
    return K


def rmsd_qcp(X, Y, GX=None, GY=None):
    """All-pairs RMSD between centred frames X (nx, a, 3) and Y (ny, a, 3), float64."""
    X = np.asarray(X, dtype=np.float64)
    Y = np.asarray(Y, dtype=np.float64)
    n_atoms = X.shape[1]
    if GX is None:
        GX = np.einsum("fad,fad->f", X, X)
    if GY is None:
        GY = np.einsum("fad,fad->f", Y, Y)
    M = np.einsum("iad,jae->ijde", X, Y)
    lam = np.linalg.eigvalsh(key_matrix(M))[..., -1]
    msd = (np.asarray(GX, dtype=np.float64)[:, None] +
           np.asarray(GY, dtype=np.float64)[None, :] - 2.0 * lam) / n_atoms
    return np.sqrt(np.maximum(msd, 0.0))


def rmsd_kabsch(X, Y):
    """Independent route: optimal proper rotation by SVD (Kabsch 1976/78)."""
    X = np.asarray(X, dtype=np.float64)
    Y = np.asarray(Y, dtype=np.float64)
    out = np.empty((len(X), len(Y)))
    for i, a in enumerate(X):
        for j, b in enumerate(Y):
            H = a.T @ b
            U, S, Vt = np.linalg.svd(H)
            d = np.sign(np.linalg.det(U @ Vt))
            R = U @ np.diag([1.0, 1.0, d]) @ Vt
            diff = a @ R - b
            out[i, j] = np.sqrt(max((diff * diff).sum() / a.shape[0], 0.0))
    return out


def dist(X, y, GX=None, Gy=None):
    """libdistance.dist(..., 'rmsd'): one-to-many (libdistance.pyx:527-562)."""
    return rmsd_qcp(X, np.asarray(y)[None] if np.asarray(y).ndim == 2 else y, GX,
                    None if Gy is None else np.atleast_1d(Gy))[:, 0]


def assign_nearest(X, Y, GX=None, GY=None):
    """libdistance.assign_nearest(..., 'rmsd') (libdistance.pyx:316-370): strict '<'."""
    D = rmsd_qcp(X, Y, GX, GY)
    labels = np.argmin(D, axis=1)  # first minimum == strict '<' scan
    return labels.astype(np.intp), float(D[np.arange(len(D)), labels].sum())


def pdist(X, X_indices=None, GX=None):
    """libdistance.pdist(..., 'rmsd') (libdistance.pyx:464-499), condensed order."""
    X = np.asarray(X)
    if X_indices is not None:
        X = X[np.asarray(X_indices)]
        if GX is not None:
            GX = np.asarray(GX)[np.asarray(X_indices)]
    D = rmsd_qcp(X, X, GX, GX)
    iu = np.triu_indices(len(X), k=1)
    return D[iu]


def cdist_rmsd(XA, XB):
    """libdistance.cdist(XA, XB, 'rmsd') on UNCENTRED coordinates (libdistance.pyx:424-440 centres
    private copies first): what RMSDFeaturizer.partial_transform returns (featurizer.py:318-320)."""
    a, ga = center_and_trace(XA)
    b, gb = center_and_trace(XB)
    return rmsd_qcp(a, b, ga, gb)
