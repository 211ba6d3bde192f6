"""CPU: the float32 filter of the fused k-centers pass (csrc/kcenters_lookahead.cu) never skips a
(frame, centre) pair that the reference comparison `d < cur` (kcenters.py:93) would accept.

The kernel's float32 arithmetic is emulated operation by operation -- float32 difference, packed
fma chains per lane (elements x,z / y,w of each float4), their sum, the xor-tree over the 32 lanes of
a group -- and compared with the reference value (float32 difference, float64 square-accumulate,
sqrt: distance_kernels.h:54-77) on adversarial inputs: current minima a few ulps around the true
distance, large common offsets, tiny and huge magnitudes, zero distances."""
import numpy as np
import pytest

F32 = np.float32


def _fma32(a, b, c):
    # a*b is exact in float64 for float32 inputs; one rounding to float32 like the hardware fma
    return F32(np.float64(a) * np.float64(b) + np.float64(c))


def kernel_filter_sum(x, c, G=32):
    """float32 squared distance exactly as kcenters_multi_pass_kernel forms it (ITERS = d/4/G)."""
    d = len(x)
    iters = d // 4 // G
    t = (x.astype(F32) + (-c.astype(F32))).astype(F32)          # FADD2 with the negated centre
    lane = np.zeros(G, dtype=F32)
    for l in range(G):
        ax, ay = F32(0), F32(0)
        for i in range(iters):
            q = t[4 * (l + i * G): 4 * (l + i * G) + 4]
            ax = _fma32(q[0], q[0], ax)
            ay = _fma32(q[1], q[1], ay)
            ax = _fma32(q[2], q[2], ax)
            ay = _fma32(q[3], q[3], ay)
        lane[l] = F32(ax + ay)
    v = lane
    while len(v) > 1:                                            # xor tree: lane l + lane l ^ (len/2)
        h = len(v) // 2
        v = (v[:h] + v[h:]).astype(F32)
    return v[0]


def reference_distance(x, c):
    df = (x.astype(F32) - c.astype(F32)).astype(F32)
    return np.sqrt(np.sum(df.astype(np.float64) ** 2))


def skipped(tot, cur, d, euclidean=True):
    one_minus_eps = F32(1.0) - F32(4.0) * F32(d + 8) * F32(5.9604645e-8)
    bound = np.float64(cur) * np.float64(cur) if euclidean else np.float64(cur)
    bound32 = np.nextafter(F32(bound), F32(np.inf)) if F32(bound) < bound else F32(bound)   # round up
    return bool(F32(tot * one_minus_eps) >= bound32 and tot >= F32(1e-30))


@pytest.mark.parametrize("d", [128, 256])
@pytest.mark.parametrize("scale,offset", [(1.0, 0.0), (1e-3, 0.0), (1e4, 0.0), (1.0, 1e3), (1.0, 1e5), (1e-6, 1.0),
                                          (1e-21, 0.0), (1e-24, 0.0)])
def test_filter_never_skips_an_accepting_pair(d, scale, offset):
    rs = np.random.RandomState(int(d + 10 * np.log10(scale + 1e-9) + offset) % 2 ** 31)
    n_skipped = 0
    for trial in range(120):
        x = (offset + scale * rs.randn(d)).astype(F32)
        c = (offset + scale * rs.randn(d)).astype(F32)
        if trial % 10 == 0:
            c = x.copy()                                         # zero distance
        if trial % 10 == 1:
            c = (x + F32(scale * 1e-4) * rs.randn(d).astype(F32)).astype(F32)   # nearly identical frames
        dref = reference_distance(x, c)
        tot = kernel_filter_sum(x, c)
        # current minima right around the true distance (where a wrong skip could happen) and far from it
        curs = [dref, np.nextafter(dref, np.inf), np.nextafter(dref, 0.0), dref * (1 + 1e-7), dref * (1 - 1e-7),
                dref * (1 + 1e-5), dref * (1 - 1e-5), dref * (1 + 1e-3), dref * 0.5, dref * 2.0, np.inf, 0.0,
                1.5e-45, 1e-30, 1e-22]
        for cur in curs:
            if skipped(tot, cur, d):
                n_skipped += 1
                assert not (dref < cur), (d, scale, offset, trial, dref, cur, tot)
    assert n_skipped > 0 or scale < 1e-15   # the filter does skip (not vacuous); denormal sums never do


def test_filter_error_is_far_inside_the_margin():
    rs = np.random.RandomState(0)
    worst = 0.0
    for _ in range(200):
        d = 256
        x = rs.randn(d).astype(F32) * F32(10 ** rs.uniform(-3, 3))
        c = rs.randn(d).astype(F32) * F32(10 ** rs.uniform(-3, 3))
        ref = reference_distance(x, c) ** 2
        tot = np.float64(kernel_filter_sum(x, c))
        worst = max(worst, abs(tot - ref) / ref)
    assert worst < 4e-6 < 4 * (256 + 8) * 2.0 ** -24            # measured error vs the margin used
