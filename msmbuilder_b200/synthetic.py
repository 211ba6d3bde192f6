"""Seeded synthetic workloads, host and device flavours.

Feature data: S sequences of L frames of D features.  Latent AR(1) processes
z_t = phi * z_{t-1} + sqrt(1 - phi^2) * eps_t with phi_i = linspace(0.999, 0.5, D)
(so the tICA eigenvalues ~ phi_i**lag are spread over (0, 1) and well separated),
mixed by a seeded orthogonal matrix and shifted by non-zero means (which
exercises the mu mu^T cancellation of tica.py:240,249):  x_t = z_t Q + mu.

The NumPy and the device generators share the statistics, not the bit pattern:
GPU benches generate on the device and hand a D2H copy of a subset to the CPU
baseline, so both arms always see identical numbers.
"""
import numpy as np


def _mixing(D, seed):
    rs = np.random.RandomState(seed + 7919)
    Q, _ = np.linalg.qr(rs.randn(D, D))
    mu = rs.uniform(-2.0, 2.0, size=D)
    phi = np.linspace(0.999, 0.5, D)
    return Q, mu, phi


def ar1_numpy(n_seq, length, D, seed=0, dtype=np.float32):
    """List of n_seq arrays (length, D)."""
    from scipy.signal import lfilter
    Q, mu, phi = _mixing(D, seed)
    rs = np.random.RandomState(seed)
    out = []
    for _ in range(n_seq):
        eps = rs.randn(length, D) * np.sqrt(1.0 - phi ** 2)
        z0 = rs.randn(D)
        z = np.empty((length, D))
        for j in range(D):
            # z_t = phi z_{t-1} + eps_t, stationary start
            zi = lfilter([1.0], [1.0, -phi[j]], eps[:, j], zi=[phi[j] * z0[j]])[0]
            z[:, j] = zi
        out.append((z @ Q + mu).astype(dtype))
    return out


def ar1_device(n_seq, length, D, seed=0, block=128, seqs_per_chunk=None, out=None, first_seq=0):
    """(n_seq * length, D) float32 CUDA tensor holding n_seq back-to-back sequences.

    `seed` fixes the process (mixing matrix, means); the noise of sequence j is drawn
    from its own generator seeded by (seed, first_seq + j), so a rank that generates
    sequences [first_seq, first_seq + n_seq) of a global dataset gets exactly the
    frames any other sharding would give it.

    Blocked linear recurrence: inside a block of `block` frames the AR(1) response
    is a (block x block) lower-triangular Toeplitz product per feature; the carry
    between blocks is a short sequential loop vectorised over sequences.
    """
    import torch
    dev = torch.device("cuda")
    Q, mu, phi = _mixing(D, seed)
    Qd = torch.from_numpy(Q).to(dev, torch.float32)
    mud = torch.from_numpy(mu).to(dev, torch.float32)
    phid = torch.from_numpy(phi).to(dev, torch.float64)
    g = torch.Generator(device=dev)
    B = block
    nb = (length + B - 1) // B
    Lp = nb * B
    # T[d, i, j] = phi_d^(i-j) for i >= j
    ar = torch.arange(B, device=dev, dtype=torch.float64)
    expo = (ar[:, None] - ar[None, :]).clamp(min=0)
    T = (phid[:, None, None] ** expo[None]) * (ar[:, None] >= ar[None, :])[None]
    T = T.to(torch.float32)                                   # (D, B, B)
    decay = (phid[None, :] ** (ar[:, None] + 1.0)).to(torch.float32)   # (B, D): phi^(i+1)
    decay_B = (phid ** B).to(torch.float32)                   # (D,)
    sig = torch.sqrt(1.0 - phid ** 2).to(torch.float32)

    if out is None:
        out = torch.empty((n_seq * length, D), dtype=torch.float32, device=dev)
    if seqs_per_chunk is None:
        seqs_per_chunk = max(1, int((1 << 28) // max(1, Lp * D)))   # ~1 GiB of float32 per temp
    for s0 in range(0, n_seq, seqs_per_chunk):
        s = min(seqs_per_chunk, n_seq - s0)
        eps = torch.empty((s, nb, B, D), device=dev, dtype=torch.float32)
        carry = torch.empty((s, D), device=dev, dtype=torch.float32)               # z_{-1}
        for j in range(s):
            g.manual_seed((int(seed) * 1000003 + int(first_seq) + s0 + j) % (2 ** 63 - 1))
            torch.randn((nb, B, D), generator=g, out=eps[j])
            torch.randn((D,), generator=g, out=carry[j])
        eps *= sig
        # within-block response: y[s, b, i, d] = sum_j T[d, i, j] eps[s, b, j, d]; one bmm per
        # sequence so that its shape (and with it the cuBLAS kernel and the bits) never
        # depends on how many sequences this call generates
        y = torch.empty((s, nb, B, D), device=dev, dtype=torch.float32)
        for j in range(s):
            e = eps[j].permute(2, 1, 0).contiguous()             # (D, B, nb)
            y[j] = torch.bmm(T, e).permute(2, 1, 0)              # (nb, B, D)
        del eps, e
        for b in range(nb):
            y[:, b] += decay[None] * carry[:, None, :]
            carry = y[:, b, B - 1, :].clone()
        for j in range(s):
            z = y[j].reshape(Lp, D)[:length]
            out[(s0 + j) * length:(s0 + j + 1) * length] = torch.matmul(z, Qd) + mud
        del y, z
    return out


def rmsd_conformations_numpy(n_frames, n_atoms=100, n_templates=20, seed=0, noise=0.05):
    """(n_frames, n_atoms, 3) float32: seeded template conformations + per-frame
    noise, a random rotation and a random translation (so superposition matters)."""
    rs = np.random.RandomState(seed)
    templates = rs.randn(n_templates, n_atoms, 3) * 0.3
    which = rs.randint(0, n_templates, size=n_frames)
    xyz = templates[which] + rs.randn(n_frames, n_atoms, 3) * noise
    # random rotations from normalised quaternions
    q = rs.randn(n_frames, 4)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    w, x, y, z = q.T
    R = np.stack([
        np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)], -1),
        np.stack([2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)], -1),
        np.stack([2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], -1)], -2)
    xyz = np.einsum("fad,fed->fae", xyz, R) + rs.randn(n_frames, 1, 3)
    return xyz.astype(np.float32), which


def rmsd_conformations_device(n_frames, n_atoms=100, n_templates=2000, seed=0, noise=0.05,
                              templates=None):
    """Device twin of rmsd_conformations_numpy.  `templates` (n_templates, n_atoms, 3): reuse one
    template bank for several calls (chunks of one data set, each with its own `seed`)."""
    import torch
    dev = torch.device("cuda")
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    if templates is None:
        templates = torch.randn((n_templates, n_atoms, 3), generator=g, device=dev) * 0.3
    else:
        n_templates = int(templates.shape[0])
    which = torch.randint(0, n_templates, (n_frames,), generator=g, device=dev)
    xyz = templates[which] + torch.randn((n_frames, n_atoms, 3), generator=g, device=dev) * noise
    q = torch.randn((n_frames, 4), generator=g, device=dev)
    q = q / q.norm(dim=1, keepdim=True)
    w, x, y, z = q.unbind(1)
    R = torch.stack([
        torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)], -1),
        torch.stack([2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)], -1),
        torch.stack([2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], -1)], -2)
    xyz = torch.einsum("fad,fed->fae", xyz, R) + torch.randn((n_frames, 1, 3), generator=g, device=dev)
    return xyz.contiguous(), which


def dihedral_standin_numpy(n_seq=10, length=9999, seed=0):
    """Stand-in for config 1 (AlanineDipeptide sin/cos of phi, psi; the real data
    needs mdtraj + network): two metastable angular coordinates from an overdamped
    double-well walk, featurised as [sin a, cos a, sin b, cos b] -> (length, 4)."""
    rs = np.random.RandomState(seed)
    out = []
    for _ in range(n_seq):
        ang = np.zeros((length, 2))
        x = rs.uniform(-1, 1, size=2)
        dt, kT = 1e-2, 0.35
        noise = rs.randn(length, 2) * np.sqrt(2 * kT * dt)
        for t in range(length):
            grad = 4 * x * (x * x - 1.0)          # V = (x^2-1)^2
            x = x - grad * dt + noise[t]
            ang[t] = x
        a, b = ang[:, 0] * 1.2, ang[:, 1] * 1.2
        out.append(np.stack([np.sin(a), np.cos(a), np.sin(b), np.cos(b)], 1))
    return out
