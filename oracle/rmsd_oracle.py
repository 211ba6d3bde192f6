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


# ---------------------------------------------------------------------------------------
# float32 restatement of the arithmetic the reference CALLS (libdistance.pyx:350-351):
#   rmsd = sqrtf(msd_atom_major(n, n, x, y, G_x, G_y, 0, NULL))
# msd_atom_major is mdtraj's libtheobald (absent here).  Its published structure (Haque, Beauchamp &
# Pande's IRMSD / mdtraj `theobald_rmsd.c`, built on Theobald's qcprot.c) is: the 3x3 inner-product
# matrix M is accumulated in FLOAT32 four atoms at a time (one SSE register per entry: a separate
# multiply and add per 4-atom group, a horizontal add at the end); the quartic's coefficients and
# the Newton iteration from E0 = (G_x + G_y) / 2 run in double (evalprec 1e-11, at most 50 steps);
# the msd comes back as a float.  The order of the float32 additions inside the SSE code cannot be
# checked without the source, so this restatement pins the PRECISION CLASS (float32 products and
# sums of M, float32 traces, float result), not the bits: parity for 'rmsd' stays unpinned at the
# bit level, and the GPU tests compare within the float32 envelope measured here.
# ---------------------------------------------------------------------------------------

# Known answer published with the method: the 7-atom fragments of Theobald's qcprot `main.c`
# (Liu, Agrafiotis & Theobald 2010, supplementary code), "QCP rmsd: 0.719106" and its rotation matrix.
QCPROT_FRAG_A = np.array([[-2.803, -15.373, 24.556], [0.893, -16.062, 25.147], [1.368, -12.371, 25.885],
                          [-1.651, -12.153, 28.177], [-0.440, -15.218, 30.068], [2.551, -13.273, 31.372],
                          [0.105, -11.330, 33.567]])
QCPROT_FRAG_B = np.array([[-14.739, -18.673, 15.040], [-12.473, -15.810, 16.074], [-14.802, -13.307, 14.408],
                          [-17.782, -14.852, 16.171], [-16.124, -14.617, 19.584], [-15.029, -11.037, 18.902],
                          [-18.577, -10.001, 17.996]])
QCPROT_RMSD = 0.719106
QCPROT_ROTATION = np.array([[0.72216358, 0.69118937, -0.0271479],
                            [-0.52038257, 0.51700833, -0.67963547],
                            [-0.45572112, 0.50493528, 0.73304748]])


def inner_products_f32_sse(x, y):
    """M[p, q] = sum_a x[a, p] * y[a, q] the way a 4-wide float32 SIMD loop forms it: lane a % 4 adds
    fl32(x * y) in atom order (separate multiply and add), then (s0 + s1) + (s2 + s3)."""
    x = np.asarray(x, dtype=np.float32)
    y = np.asarray(y, dtype=np.float32)
    n = x.shape[0]
    lanes = np.zeros((4, 3, 3), dtype=np.float32)
    for a in range(n):
        prod = (x[a][:, None] * y[a][None, :]).astype(np.float32)
        lanes[a & 3] = (lanes[a & 3] + prod).astype(np.float32)
    return ((lanes[0] + lanes[1]).astype(np.float32) + (lanes[2] + lanes[3]).astype(np.float32)).astype(np.float32)


def _det3(a, b, c, d, e, f, g, h, i):
    # (a (e i - f h) - b (d i - f g)) + c (d h - e g): plain Python floats, one IEEE rounding per operation
    return (a * (e * i - f * h) - b * (d * i - f * g)) + c * (d * h - e * g)


def msd_from_M_and_G(M, Ga, Gb, n_atoms, evalprec=1e-11):
    """Published QCP Newton iteration (qcprot.c FastCalcRMSDAndRotation) in double on a given M.
    Written operation by operation (no fused multiply-add, fixed order) so that
    msmbuilder_b200/csrc/rmsd.cu:qcp_rmsd_strict can be compared with it bit for bit."""
    m = [float(v) for v in np.asarray(M, dtype=np.float64).reshape(9)]
    Sxx, Sxy, Sxz, Syx, Syy, Syz, Szx, Szy, Szz = m
    fro = 0.0
    for v in m:
        fro = fro + v * v
    c2 = -2.0 * fro
    c1 = -8.0 * _det3(Sxx, Sxy, Sxz, Syx, Syy, Syz, Szx, Szy, Szz)
    k00 = (Sxx + Syy) + Szz
    k01 = Syz - Szy
    k02 = Szx - Sxz
    k03 = Sxy - Syx
    k11 = (Sxx - Syy) - Szz
    k12 = Sxy + Syx
    k13 = Szx + Sxz
    k22 = (Syy - Sxx) - Szz
    k23 = Syz + Szy
    k33 = Szz - (Sxx + Syy)
    d0 = _det3(k11, k12, k13, k12, k22, k23, k13, k23, k33)
    d1 = _det3(k01, k12, k13, k02, k22, k23, k03, k23, k33)
    d2 = _det3(k01, k11, k13, k02, k12, k23, k03, k13, k33)
    d3 = _det3(k01, k11, k12, k02, k12, k22, k03, k13, k23)
    c0 = ((k00 * d0 - k01 * d1) + k02 * d2) - k03 * d3
    e0 = 0.5 * (float(Ga) + float(Gb))
    lam = e0
    for _ in range(50):
        old = lam
        x2 = lam * lam
        b = (x2 + c2) * lam
        a = b + c1
        num = a * lam + c0
        den = ((2.0 * x2) * lam + b) + a
        lam = lam - num / den
        if abs(lam - old) < abs(evalprec * lam):
            break
    return abs((2.0 * (e0 - lam)) / float(n_atoms))


def rmsd_theobald_f32(X, Y, GX, GY):
    """All pairs, float32 accumulation of M (SIMD order), double Newton, float msd, float sqrt."""
    X = np.asarray(X, dtype=np.float32)
    Y = np.asarray(Y, dtype=np.float32)
    out = np.empty((len(X), len(Y)), dtype=np.float32)
    for i in range(len(X)):
        for j in range(len(Y)):
            msd = np.float32(msd_from_M_and_G(inner_products_f32_sse(X[i], Y[j]), np.float32(GX[i]),
                                              np.float32(GY[j]), X.shape[1]))
            out[i, j] = np.sqrt(msd, dtype=np.float32)
    return out
