"""tools/config_checks.py -- the BASELINE.json parity configs (2, 3, 5) at sizes that finish in
about a minute on one B200, each with its parity statement and device timings."""
import os
import sys
import time
import warnings

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from msmbuilder_b200 import _kernels as K
from msmbuilder_b200.cluster import KCenters
from msmbuilder_b200.decomposition import tICA
from msmbuilder_b200.synthetic import ar1_device, rmsd_conformations_device
from oracle.tica_oracle import TicaOracle
from oracle import libdistance_oracle as lo
from oracle import rmsd_oracle as ro


def timed(fn):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    out = fn()
    b.record()
    torch.cuda.synchronize()
    return out, a.elapsed_time(b)


def config2(n_seq=100, L=100000, D=64):
    X = ar1_device(n_seq, L, D, seed=2)
    seqs = [X[i * L:(i + 1) * L] for i in range(n_seq)]
    m = tICA(n_components=4, lag_time=10)
    m.fit(seqs)
    m2 = tICA(n_components=4, lag_time=10)
    m2._initialize(D)
    _, ms = timed(lambda: m2._accumulate_device(seqs))
    host = [s.cpu().numpy() for s in seqs[:20]]
    t0 = time.time()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        o = TicaOracle(n_components=4, lag_time=10).fit(host)
        g = tICA(n_components=4, lag_time=10).fit(host)
    t_cpu = time.time() - t0
    print("config 2: tICA fit %d x %d f32: %.2f ms on device (%.1f Mframes/s, engine auto); "
          "eigenvalues vs f64 oracle on a %d-frame prefix: max |d| = %.2e (tol 1e-5); oracle+gpu prefix %.1f s"
          % (n_seq * L, D, ms, n_seq * L / ms / 1e3, 20 * L, np.abs(g.eigenvalues_ - o.eigenvalues_).max(), t_cpu))


def config3(n=10_000_000, D=16, k=500):
    g = torch.Generator(device="cuda"); g.manual_seed(3)
    X = torch.randn((n, D), generator=g, device="cuda") * torch.linspace(3, 0.3, D, device="cuda")
    C = X[torch.randint(0, n, (k,), generator=g, device="cuda")].contiguous()
    (labels, dmin, inertia), ms = timed(lambda: K.assign_nearest(X, C, "euclidean", want_min_dist=True))
    (labels, dmin, inertia), ms = timed(lambda: K.assign_nearest(X, C, "euclidean", want_min_dist=True))
    os.environ["MSMB200_ASSIGN_EXACT"] = "1"
    (l2, d2, i2), ms_exact = timed(lambda: K.assign_nearest(X[:2_000_000], C, "euclidean", want_min_dist=True))
    os.environ.pop("MSMB200_ASSIGN_EXACT")
    same = bool((labels[:2_000_000] == l2).all())
    ns = 200_000
    t0 = time.time()
    ref, ref_inertia = lo.assign_nearest(X[:ns].cpu().numpy(), C.cpu().numpy(), "euclidean",
                                         impl="reference" if lo.have_reference() else "port")
    t_cpu = time.time() - t0
    ok = np.array_equal(labels[:ns].cpu().numpy(), ref)
    print("config 3: assign %d x %d to k=%d: %.2f ms (%.1f Mframes/s) fast engine; exact engine %.2f ms per 2M; "
          "fast == exact labels on 2M: %s; == reference C++ on %d frames: %s (reference: %.2f s, %.3f Mframes/s)"
          % (n, D, k, ms, n / ms / 1e3, ms_exact, same, ns, ok, t_cpu, ns / t_cpu / 1e6))


def config5(n=500_000, n_atoms=100, k=100):
    xyz, which = rmsd_conformations_device(n, n_atoms, n_templates=k, seed=5)
    kc = KCenters(n_clusters=k, metric="rmsd", random_state=0)
    _, ms = timed(lambda: kc.fit([xyz]))
    c, G = ro.center_and_trace(xyz[:20000].cpu().numpy())
    cent, Gc = ro.center_and_trace(kc.cluster_centers_)
    D = ro.rmsd_qcp(c, cent, G, Gc)
    lab = np.concatenate(kc.labels_)[:20000]
    Ds = np.sort(D, axis=1)
    clear = (Ds[:, 1] - Ds[:, 0]) > 1e-4
    agree = (D.argmin(1)[clear] == lab[clear]).mean()
    print("config 5 (scaled): KCenters(k=%d, rmsd) on %d x %d atoms: %.1f ms incl. D2H (%.2f ms/pass); "
          "%d distinct templates among the %d centres; labels vs f64 QCP oracle on 20k frames (clear margins): %.4f"
          % (k, n, n_atoms, ms, ms / k, len(set(which[kc.cluster_ids_].tolist())), k, agree))


if __name__ == "__main__":
    config2()
    config3()
    config5()
