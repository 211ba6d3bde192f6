"""K3 timing points on the GPU box (CUDA events, best of 3): assign_nearest for float32 euclidean at the
shapes the tensor-core filters take, with the engine each one uses and a float64-scan label check.

    python tools/assign_points.py [n,d,k ...]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

ENGINES = {0: "tcgen05 resident", 1: "tcgen05 streamed", 2: "SIMT"}


def main():
    import torch
    from msmbuilder_b200 import _kernels as K, _lib
    lib = _lib.load()
    points = [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]] or [
        (10_000_000, 16, 500), (10_000_000, 128, 2000), (10_000_000, 128, 256), (10_000_000, 64, 1000),
        (10_000_000, 256, 8), (4_000_000, 256, 1000)]
    for n, d, k in points:
        g = torch.Generator(device="cuda")
        g.manual_seed(n % 1000 + d + k)
        X = torch.randn((n, d), generator=g, device="cuda") * torch.linspace(3, 0.3, d, device="cuda")
        C = X[torch.randint(0, n, (k,), generator=g, device="cuda")].contiguous()

        def run(x=X):
            return K.assign_nearest(x, C, "euclidean")

        def best(fn, reps=3):
            fn()
            torch.cuda.synchronize()
            ts = []
            for _ in range(reps):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                r = fn()
                e1.record()
                e1.synchronize()
                ts.append(e0.elapsed_time(e1))
            return min(ts), r

        ms, (labels, _, _) = best(run)
        m = min(n, 500_000)
        os.environ["MSMB200_ASSIGN_EXACT"] = "1"
        exact, _, _ = K.assign_nearest(X[:m], C, "euclidean")
        os.environ.pop("MSMB200_ASSIGN_EXACT")
        os.environ["MSMB200_ASSIGN_SIMT"] = "1"
        ms_simt, _ = best(lambda: run(X[:m]), reps=1)
        os.environ.pop("MSMB200_ASSIGN_SIMT")
        print("n=%d d=%d k=%d  %-17s %8.3f ms  %7.1f M frames/s  %6.1f TFLOP/s alg  labels==f64 scan on %d: %s"
              "  | SIMT filter: %.1f ms per %d frames"
              % (n, d, k, ENGINES[int(lib.msmb200_assign_engine(n, k, d))], ms, n / ms / 1e3,
                 2.0 * n * k * d / ms / 1e9, m, bool((labels[:m] == exact).all()), ms_simt * n / m, n), flush=True)
        del X, C, labels, exact
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
