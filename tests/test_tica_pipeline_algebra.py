"""CPU: the algebra of the tensor-core tICA path (csrc/tica_umma.cu) restated in NumPy and checked
against the oracle of tica.py:401-424.

What the device does is NOT the reference's six products on raw frames: it centres every frame by a
provisional float32 mean, scales it by a per-feature power of two, splits it into two fp16 parts,
sums three products over the pair rows, takes the 2*lag head/tail rows in float64, derives C_tautau
and S_tau from them, and undoes shift and scale in float64 (tica_umma_finalize_kernel).  This test
follows those steps with exact float64 sums in place of the tensor cores and must land on the
reference's raw moments -- so a formula slip in shift / scale / head / tail handling shows up on
the CPU suite, independently of the GPU parity tests."""
import warnings

import numpy as np
import pytest

from msmbuilder_b200.synthetic import ar1_numpy
from oracle.tica_oracle import TicaOracle

F32 = np.float32


def _round_to_bits(a, bits=11):
    """(bits + half) & mask on the float32 pattern: the kernel's integer rounding of the high part."""
    u = a.astype(F32).view(np.uint32)
    add = np.uint32(1 << (23 - bits))
    mask = np.uint32(~((1 << (24 - bits)) - 1) & 0xFFFFFFFF)
    return ((u + add) & mask).view(F32)


def device_algebra(seqs, lag, split=True, all_rows=False):
    """all_rows: the MN-major mode of tica_umma_v2.cuh -- G = sum over EVERY row of h (h/2)^T + h l^T (C_00 = G + G^T)
    and the column sums over every row; the finalize kernel takes the tail rows out of C_00 / S_0 and the head
    rows out of C_tautau / S_tau."""
    D = seqs[0].shape[1]
    usable = [np.ascontiguousarray(s, dtype=F32) for s in seqs if len(s) > lag]
    allrows = np.concatenate(usable)                                             # tica_shift_kernel: 1024 rows
    rows = min(len(allrows), 1024)                                               # spread over the whole call
    first = allrows[(np.arange(rows, dtype=np.int64) * len(allrows)) // rows]
    shift = (first.astype(np.float64).sum(0) / len(first)).astype(F32)
    m = np.maximum(np.abs(first - shift).max(0), np.abs(shift) / F32(256)).astype(F32)
    e = np.where((m > 0) & np.isfinite(m), np.floor(np.log2(np.maximum(m, 1e-300))), 0)
    scale = np.ldexp(F32(1), -e.astype(int)).astype(F32)
    Ctau = np.zeros((D, D)); C00 = np.zeros((D, D)); S0 = np.zeros(D)
    E2 = np.zeros((D, D)); E3 = np.zeros((D, D)); head = np.zeros(D); tail = np.zeros(D)
    n_pairs = n_obs = 0
    for X in usable:
        n = len(X)
        xp = (X - shift).astype(F32)                                             # x' = fl32(x - shift)
        a = (xp * scale).astype(F32)                                             # exact: power of two
        if split:                                                                # fp16 h + l, three products
            h = _round_to_bits(a)
            l = (a - h).astype(F32).astype(np.float16).astype(np.float64)
            h = h.astype(np.float16).astype(np.float64)
            A, B = (h[:n - lag], l[:n - lag]), (h[lag:], l[lag:])
            Ctau += A[0].T @ B[0] + A[0].T @ B[1] + A[1].T @ B[0]
            if all_rows:
                G = h.T @ (h / 2) + h.T @ l
                C00 += G + G.T
            else:
                C00 += A[0].T @ A[0] + A[0].T @ A[1] + A[1].T @ A[0]
        else:
            ad = a.astype(np.float64)
            Ctau += ad[:n - lag].T @ ad[lag:]
            C00 += ad.T @ ad if all_rows else ad[:n - lag].T @ ad[:n - lag]
        S0 += (a if all_rows else a[:n - lag]).astype(np.float64).sum(0)         # column sums, scaled units
        xd = xp.astype(np.float64)                                               # edge kernel: float64, unscaled
        E2 += xd[:lag].T @ xd[:lag]
        E3 += xd[n - lag:].T @ xd[n - lag:]
        head += xd[:lag].sum(0)
        tail += xd[n - lag:].sum(0)
        n_pairs += n - lag
        n_obs += n
    inv = 1.0 / scale.astype(np.float64)                                          # finalize
    Ctau *= np.outer(inv, inv)
    C00 *= np.outer(inv, inv)
    S0 = S0 * inv
    if all_rows:
        Ctt = C00 - E2
        C00 = C00 - E3
        St = S0 - head
        S0 = S0 - tail
    else:
        Ctt = C00 - E2 + E3
        St = S0 - head + tail
    sh = shift.astype(np.float64)
    Np = float(n_pairs)
    raw_tau = Ctau + np.outer(S0, sh) + np.outer(sh, St) + Np * np.outer(sh, sh)
    raw_00 = C00 + np.outer(S0, sh) + np.outer(sh, S0) + Np * np.outer(sh, sh)
    raw_tt = Ctt + np.outer(St, sh) + np.outer(sh, St) + Np * np.outer(sh, sh)
    return np.concatenate([raw_tau.ravel(), raw_00.ravel(), raw_tt.ravel(), S0 + Np * sh, St + Np * sh,
                           S0 + tail + n_obs * sh, [float(n_obs), float(len(usable))]])


@pytest.mark.parametrize("lag", [1, 7, 10])
@pytest.mark.parametrize("split", [False, True])
@pytest.mark.parametrize("all_rows", [False, True])
def test_device_algebra_reproduces_the_reference_moments(lag, split, all_rows):
    lens = [1500, 700, 64, lag + 1, lag, 3, 900]                 # the last three: barely usable / skipped
    seqs = [s[:n] for s, n in zip(ar1_numpy(len(lens), 1500, 32, seed=5), lens)]
    rs = np.random.RandomState(1)
    seqs = [(s * 10.0 ** rs.uniform(-3, 3, size=32) + rs.uniform(-50, 50, size=32)).astype(F32) for s in seqs]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = TicaOracle(n_components=3, lag_time=lag).fit(seqs).packed_moments()
    got = device_algebra(seqs, lag, split=split, all_rows=all_rows)
    D = 32
    assert got[-2] == ref[-2] and got[-1] == ref[-1]             # n_observations_, n_sequences_
    sd = np.sqrt(np.abs(np.diag(ref[D * D:2 * D * D].reshape(D, D))))
    norm = np.outer(sd, sd)
    for k in range(3):                                           # C_tau, C_00, C_tautau per unit of scale
        a = got[k * D * D:(k + 1) * D * D].reshape(D, D) / norm
        b = ref[k * D * D:(k + 1) * D * D].reshape(D, D) / norm
        assert np.abs(a - b).max() < (2e-6 if split else 1e-6), k
    for k in range(3):                                           # S_0, S_tau, S
        a, b = got[3 * D * D + k * D:3 * D * D + (k + 1) * D], ref[3 * D * D + k * D:3 * D * D + (k + 1) * D]
        np.testing.assert_allclose(a, b, rtol=1e-6, atol=1e-6 * np.abs(b).max())
