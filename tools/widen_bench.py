"""Device timings of the section-8(f) rows (not part of bench.py's JSON line):
transition counting on device labels, RegularSpatial, KMedoids' pdist.

    python tools/widen_bench.py            # on a GPU box
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from msmbuilder_b200 import _lib, _device as dev, _kernels as K          # noqa: E402
from msmbuilder_b200.msm import transition_counts                          # noqa: E402
from msmbuilder_b200.synthetic import ar1_device                           # noqa: E402


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def counts_case(n, n_states, lag, sticky):
    g = torch.Generator(device="cuda").manual_seed(1)
    y = torch.randint(0, n_states, (n,), generator=g, device="cuda", dtype=torch.int32)
    if sticky:
        # metastable chains: hold each state for ~64 frames (the contended diagonal)
        y = y[::64].repeat_interleave(64)[:n].contiguous()
    offsets = torch.arange(0, n + 1, 100000, device="cuda", dtype=torch.int64)
    if int(offsets[-1]) != n:
        offsets = torch.cat([offsets, torch.tensor([n], device="cuda")])
    counts = torch.zeros((n_states, n_states), dtype=torch.int64, device="cuda")

    def run():
        _lib.call("msmb200_transition_counts", dev.ptr(y), 4, dev.ptr(offsets),
                  int(offsets.numel() - 1), n, lag, None, 0, 0, n_states, dev.ptr(counts),
                  dev.stream_ptr())
    ms = timed(run)
    # end to end through the Python function (range + presence + counts + D2H of the matrix)
    seqs = [y[i:i + 100000] for i in range(0, n, 100000)]
    t0 = time.perf_counter()
    c, m = transition_counts(seqs, lag)
    torch.cuda.synchronize()
    e2e = (time.perf_counter() - t0) * 1e3
    print("transition_counts n=%d n_states=%d lag=%d sticky=%d: kernel %.3f ms (%.1f G labels/s, "
          "%.0f GB/s of 4-byte labels), python call %.1f ms, total counts %.0f"
          % (n, n_states, lag, sticky, ms, n / ms / 1e6, 4.0 * n / ms / 1e6, e2e, c.sum() * lag))


def regular_spatial_case(n, d, target_centres):
    X = ar1_device(max(n // 100000, 1), min(n, 100000), d, seed=0)
    d0 = K.dist(X, X[0].contiguous(), "euclidean")
    q = torch.quantile(d0[1:200000].float(), 0.5).item()
    for scale in (1.0, 0.9, 0.8):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ids = K.regular_spatial_fit(X, q * scale, "euclidean")
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print("RegularSpatial n=%d d=%d d_min=%.3f: %d centres in %.1f ms (%.3f ms per centre, "
              "%.1f GB/s of frames after the centre on average)"
              % (n, d, q * scale, len(ids), dt * 1e3, dt * 1e3 / len(ids),
                 0.5 * n * d * 4 * len(ids) / dt / 1e9))
        if len(ids) >= target_centres:
            break


def kmedoids_case(n, d, k, n_passes):
    from msmbuilder_b200.cluster import KMedoids
    X = ar1_device(1, n, d, seed=2)
    ms = timed(lambda: K.pdist(X, "euclidean"), reps=3)
    pairs = n * (n - 1) // 2
    t0 = time.perf_counter()
    km = KMedoids(n_clusters=k, n_passes=n_passes, random_state=0).fit([X])
    dt = time.perf_counter() - t0
    print("KMedoids n=%d d=%d k=%d n_passes=%d: pdist %.2f ms (%.2f G pairs/s), fit %.1f ms, inertia %.3f"
          % (n, d, k, n_passes, ms, pairs / ms / 1e6, dt * 1e3, km.inertia_))


def stream_case(n_files, length, d):
    """tICA.fit on a directory of .npy files: list of np.load arrays vs NumpyDirStream."""
    import shutil
    import tempfile
    from msmbuilder_b200.decomposition import tICA
    from msmbuilder_b200.io import NumpyDirStream, save_sequences
    tmp = tempfile.mkdtemp(prefix="msmb200_ds_")
    try:
        X = ar1_device(n_files, length, d, seed=5)
        save_sequences(tmp, [X[i * length:(i + 1) * length].cpu().numpy() for i in range(n_files)])
        del X
        nbytes = n_files * length * d * 4
        stream = NumpyDirStream(tmp, prefetch=2)
        for label, make in (("np.load list", lambda: [np.load(f) for f in stream.files]),
                            ("NumpyDirStream", lambda: stream)):
            best = 1e9
            for _ in range(3):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                m = tICA(n_components=4, lag_time=10).fit(make())
                torch.cuda.synchronize()
                best = min(best, time.perf_counter() - t0)
            print("tICA.fit %d x %d x %d from .npy (page cache), %s: %.1f ms = %.2f GB/s, %.1f M frames/s, "
                  "lambda0 %.6f" % (n_files, length, d, label, best * 1e3, nbytes / best / 1e9,
                                    n_files * length / best / 1e6, m.eigenvalues_[0]))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    _lib.require_gpu()
    for n_states, sticky in ((8, 1), (8, 0), (90, 1), (500, 1), (2000, 1), (2000, 0)):
        counts_case(50_000_000, n_states, 10, sticky)
    regular_spatial_case(2_000_000, 64, 50)
    kmedoids_case(8000, 64, 10, 3)
    stream_case(40, 100000, 256)
