"""Estimator-level multi-GPU from ONE process (`devices=`): tICA.fit and KCenters.fit on pageable
NumPy arrays, one GPU against every GPU of the box.  Prints one JSON line."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from msmbuilder_b200.decomposition import tICA
    from msmbuilder_b200.cluster import KCenters
    n_dev = torch.cuda.device_count()
    n_seq, L, D = int(os.environ.get("E2E_SEQS", 160)), 100_000, 256
    rs = np.random.RandomState(0)
    base = rs.randn(L + 64, D).astype(np.float32)
    # AR(1)-like sequences without a long generation phase: shifted windows of one noise block, mixed
    seqs = []
    for i in range(n_seq):
        s = base[(i * 7) % 64:(i * 7) % 64 + L].copy()
        s[1:] += 0.9 * s[:-1]
        s += np.float32(0.01 * i)
        seqs.append(s)
    out = {"frames": n_seq * L, "features": D, "gpus": n_dev,
           "host_memory": "pageable NumPy arrays, one per sequence"}

    def timed(fn):
        fn()                                  # warm-up: pinned rings, workspaces, tensor maps
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = fn()
        for d in range(n_dev):
            torch.cuda.synchronize(d)
        return time.perf_counter() - t0, r

    t1, a = timed(lambda: tICA(n_components=4, lag_time=10).fit(seqs))
    tn, b = timed(lambda: tICA(n_components=4, lag_time=10, devices="all").fit(seqs))
    out["tica_fit_s"] = {"1": t1, str(n_dev): tn}
    out["tica_eig_diff"] = float(np.abs(a.eigenvalues_ - b.eigenvalues_).max())
    k1, c = timed(lambda: KCenters(n_clusters=8, random_state=0).fit(seqs))
    kn, d = timed(lambda: KCenters(n_clusters=8, random_state=0, devices="all").fit(seqs))
    out["kcenters_fit_s"] = {"1": k1, str(n_dev): kn}
    out["kcenters_ids_equal"] = c.cluster_ids_ == d.cluster_ids_
    out["kcenters_labels_equal"] = all(np.array_equal(x, y) for x, y in zip(c.labels_, d.labels_))
    out["frames_per_s_both_fits"] = {"1": n_seq * L / (t1 + k1), str(n_dev): n_seq * L / (tn + kn)}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
