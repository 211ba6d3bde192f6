"""tools/e2e_probe.py -- where the end-to-end (host arrays in, results out) time goes:
raw pinned H2D rate of the box vs the two estimator calls of bench.py's e2e leg."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from msmbuilder_b200.decomposition import tICA
from msmbuilder_b200.cluster import KCenters
from msmbuilder_b200.synthetic import ar1_device

L, D, S = 100000, 256, 40
X = ar1_device(S, L, D, seed=1000)
host = [X[i * L:(i + 1) * L].cpu().pin_memory() for i in range(S)]
dst = torch.empty_like(X)
torch.cuda.synchronize()
for rep in range(3):
    t0 = time.perf_counter()
    for i, h in enumerate(host):
        dst[i * L:(i + 1) * L].copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("raw pinned H2D: %.1f ms for %.2f GB = %.1f GB/s" % (dt * 1e3, X.numel() * 4 / 1e9, X.numel() * 4 / dt / 1e9))
big = X.cpu().pin_memory()
torch.cuda.synchronize()
t0 = time.perf_counter(); dst.copy_(big, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("raw pinned H2D, one 4.1 GB copy: %.1f ms = %.1f GB/s" % (dt * 1e3, X.numel() * 4 / dt / 1e9))
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    t = tICA(n_components=4, lag_time=10).fit(host)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    kc = KCenters(n_clusters=8, random_state=0).fit(host)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print("tICA.fit(host) %.1f ms | KCenters.fit(host) %.1f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
kc = KCenters(n_clusters=8, random_state=0).fit(host)
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
