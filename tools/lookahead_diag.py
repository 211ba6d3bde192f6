"""Why does a look-ahead chain stop?  On the bench data set (AR(1), 256 features): the candidate set after the first
pass (count, tau, where tau comes from), the picks of the chain, and the number of reads of the frames per candidate
cap T.

    python tools/lookahead_diag.py [--frames 50000000] [--k 8] [--caps 512,1024,2048,4096]
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=50_000_000)
    ap.add_argument("--features", type=int, default=256)
    ap.add_argument("--k", type=int, default=8)
    ap.add_argument("--caps", default="512,1024,2048,4096")
    a = ap.parse_args()
    import torch
    from msmbuilder_b200 import _kernels as K
    from msmbuilder_b200.synthetic import ar1_device
    L = 100_000
    X = ar1_device(a.frames // L, L, a.features, seed=1000)
    n = X.shape[0]
    seed = 12345 % n
    ref_ids = None
    for cap in [int(c) for c in a.caps.split(",")]:
        st = K.LookaheadState(X, "euclidean", t_cap=cap)
        stats = {"time_passes": True}
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        K.kcenters_fit_lookahead(X, a.k, "euclidean", seed, state=K.LookaheadState(X, "euclidean", t_cap=cap))
        torch.cuda.synchronize()
        e0.record()
        ids, rows, dist, lab = K.kcenters_fit_lookahead(X, a.k, "euclidean", seed, stats=stats, state=st)
        e1.record()
        e1.synchronize()
        ids = ids.cpu().numpy()
        if ref_ids is None:
            ref_ids = ids
        passes = [(nc, round(x.elapsed_time(y), 2)) for nc, x, y in stats["pass_events"]]
        print("T = %5d: %.2f ms, %d reads %s  ids equal T0: %s" % (cap, e0.elapsed_time(e1), stats["passes"], passes,
                                                                 bool((ids == ref_ids).all())), flush=True)
    # anatomy of the first candidate set at the default cap
    st = K.LookaheadState(X, "euclidean")
    st.seed(seed)
    st.multi_pass(1, 0, first=True)
    cs = st.select().cpu()
    count = int(cs[:4].view(torch.int32)[0])
    tau = float(cs[8:16].view(torch.float64)[0])
    vals = cs[32:32 + 8 * st.t_cap].view(torch.float64)[:count].numpy()
    lane = st.lane.cpu()
    n_slots = int(lane[:8].view(torch.int64)[0])
    rec = lane[32:32 + 48 * n_slots].view(torch.float64).reshape(n_slots, 6).numpy()   # v1 i1 v2 i2 v3 pad
    v1, v2, v3 = rec[:, 0], rec[:, 2], rec[:, 4]
    sv = np.sort(vals)[::-1]
    d_all = np.sort(st.distances.cpu().numpy())[::-1]
    ent = np.sort(np.concatenate([v1, v2]))[::-1]
    print("first set: count %d  tau %.6f  | largest lane third best %.6f  T-th largest entry %.6f  (largest runner-up %.6f)"
          % (count, tau, v3.max(), ent[min(st.t_cap, len(ent)) - 1], v2.max()))
    print("top candidate values:", np.round(sv[:12], 4))
    print("rank of tau among ALL frames' distances: %d" % int((d_all > tau).sum()))
    print("lane slots %d; top-20 frames' distances: %s" % (n_slots, np.round(d_all[:20], 4)))
    # replay the chain on the host: which pick fails, and how far below tau it is
    ids_t, _, _, _ = K.kcenters_fit_lookahead(X, a.k, "euclidean", seed)
    cen = X[ids_t.long()].double()
    dm = torch.cdist(cen, cen).cpu().numpy()
    print("centre ids", ids_t.cpu().numpy())
    print("pairwise centre distances (min off-diagonal per centre):", np.round(np.where(np.eye(len(dm)) > 0, np.inf, dm).min(1), 3))
    cur = torch.full((n,), float("inf"), dtype=torch.float64, device="cuda")
    for j in range(a.k):
        for c0 in range(0, n, 1_000_000):
            blk = X[c0:c0 + 1_000_000].double()
            dj = (blk - cen[j]).pow(2).sum(1).sqrt()
            cur[c0:c0 + 1_000_000] = torch.minimum(cur[c0:c0 + 1_000_000], dj)
        top = torch.topk(cur, 3).values.cpu().numpy()
        print("after centre %d: three largest running minima %s  (tau of the first set %.6f)" % (j, np.round(top, 5), tau))

if __name__ == "__main__":
    main()
