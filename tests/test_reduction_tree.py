"""CPU: the split reduction of the distance kernels (common.cuh: group_reduce_split, and its float
twin in kcenters_lookahead.cu) adds in exactly the tree of the plain xor butterfly
(group_combine), so R group sums cost ~one reduction and stay bit-identical -- the property the
"K2b == K2 bit for bit" parity rests on.  Both are emulated lane by lane in NumPy float64/float32."""
import numpy as np
import pytest


def butterfly(vals):
    """group_combine: every lane ends with the sum; v[l] += v[l ^ off] for off = G/2 ... 1."""
    v = vals.copy()
    G = len(v)
    off = G // 2
    while off:
        v = np.array([v[l] + v[l ^ off] for l in range(G)], dtype=v.dtype)
        off //= 2
    return v


def split_reduce(vals_per_lane):
    """group_reduce_split: vals_per_lane[l] = the V values lane l holds; returns, per lane, (index of
    the value it ends up with, its total)."""
    G, V = vals_per_lane.shape
    v = [list(vals_per_lane[l]) for l in range(G)]
    off, m = G // 2, V
    while m > 1:
        new = []
        for l in range(G):
            up = (l & off) != 0
            keep = v[l][m // 2:m] if up else v[l][:m // 2]
            partner = v[l ^ off]
            recv = partner[m // 2:m] if up else partner[:m // 2]      # partner sends the half this lane keeps
            new.append([k + r for k, r in zip(keep, recv)])
        v = new
        m //= 2
        off //= 2
    r = np.array([x[0] for x in v], dtype=vals_per_lane.dtype)
    while off:
        r = np.array([r[l] + r[l ^ off] for l in range(G)], dtype=r.dtype)
        off //= 2
    which = [l // (G // V) for l in range(G)]
    return which, r


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("G,V", [(32, 4), (32, 16), (32, 8), (32, 2), (16, 4), (16, 16), (8, 4), (8, 8), (4, 4), (4, 1)])
def test_split_reduction_is_the_butterfly_tree(G, V, dtype):
    rs = np.random.RandomState(G * 100 + V)
    for _ in range(20):
        vals = (rs.randn(G, V) * 10.0 ** rs.uniform(-6, 6, size=(G, V))).astype(dtype)
        which, tot = split_reduce(vals)
        for q in range(V):
            ref = butterfly(vals[:, q].copy())
            lanes = [l for l in range(G) if which[l] == q]
            assert len(lanes) == G // V
            for l in lanes:
                # bit-identical: floating-point addition is commutative, and the pairing is the same tree
                assert tot[l] == ref[l] or (np.isnan(tot[l]) and np.isnan(ref[l]))
