"""CPU: the candidate rule of assign_refine_wide_kernel (csrc/dist_kernels.cu) never drops the centre the exact scan
(assign.hpp:50-91: float32 difference, float64 square-accumulate, lowest index on ties) would pick.

The kernel scans all k centres in float32 (lane = centre: a sequential fma chain over the features), takes the smallest
value m, and recomputes in float64 -- in ascending centre order with strict '<' -- only the centres whose float32 value
is <= max(m (1 + 16 (d + 4) 2^-24), 1e-30).  The float32 chain is emulated operation by operation and the rule is checked
on adversarial inputs: near ties a few ulps apart, exact ties / duplicate centres, common offsets, tiny and huge scales."""
import numpy as np
import pytest

F32 = np.float32


def f32_scan(x, C):
    """acc = fmaf(t, t, acc) over the features, t = float32(x - c): one float32 value per centre."""
    t = (x[None, :].astype(F32) - C.astype(F32)).astype(F32)
    acc = np.zeros(len(C), dtype=F32)
    for e in range(t.shape[1]):
        acc = (t[:, e].astype(np.float64) * t[:, e].astype(np.float64) + acc.astype(np.float64)).astype(F32)
    return acc


def exact_label(x, C):
    t = (x[None, :].astype(F32) - C.astype(F32)).astype(F32).astype(np.float64)
    d2 = np.zeros(len(C))
    for e in range(t.shape[1]):
        d2 = d2 + t[:, e] * t[:, e]
    return int(np.argmin(np.sqrt(d2)))                     # first index among equals, like the strict '<' scan


def candidates(s, d):
    rel = F32(4.0) * (F32(4.0) * F32(d + 4) * F32(5.9604645e-8))
    m = s.min()
    thr = max(F32(m * (F32(1.0) + rel)), F32(1e-30))
    return np.nonzero(s <= thr)[0]


@pytest.mark.parametrize("d", [16, 128, 256])
@pytest.mark.parametrize("scale,offset", [(1.0, 0.0), (1e-3, 0.0), (1e4, 0.0), (1.0, 1e3), (1.0, 1e5), (1e-18, 0.0),
                                          (1e-22, 0.0)])
def test_candidate_rule_keeps_the_exact_winner(d, scale, offset):
    rs = np.random.RandomState(d + int(offset) % 1000 + int(-np.log10(scale)))
    k = 96
    widest = 0
    for trial in range(40):
        C = (rs.randn(k, d) * scale + offset).astype(F32)
        x = (rs.randn(d) * scale + offset).astype(F32)
        kind = trial % 4
        if kind == 1:                                      # two centres a few ulps apart around the winner
            j = exact_label(x, C)
            C[(j + 7) % k] = C[j]
            C[(j + 7) % k, rs.randint(d)] = np.nextafter(C[j, 0], F32(np.inf))
        elif kind == 2:                                    # exact duplicates: the lowest index must stay in
            j = exact_label(x, C)
            C[(j + 11) % k] = C[j]
            C[(j + 50) % k] = C[j]
        elif kind == 3:                                    # the frame IS a centre (distance 0) next to near copies
            C[5] = x
            C[60] = x
            C[61] = x
            C[61, 0] = np.nextafter(x[0], F32(np.inf))
        s = f32_scan(x, C)
        cand = candidates(s, d)
        assert exact_label(x, C) in cand, (d, scale, offset, trial)
        widest = max(widest, len(cand))
    assert widest <= k
