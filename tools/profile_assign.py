"""One assign_nearest call at BASELINE.json config 3 (10M x 16, k = 500) for ncu / timing.

    python tools/profile_assign.py [n d k]      e.g. 10000000 128 2000 (the streamed-centres kernel)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from msmbuilder_b200 import _kernels as K

n, D, k = (int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (10_000_000, 16, 500)
g = torch.Generator(device="cuda")
g.manual_seed(3)
X = torch.randn((n, D), generator=g, device="cuda") * torch.linspace(3, 0.3, D, device="cuda")
C = X[torch.randint(0, n, (k,), generator=g, device="cuda")].contiguous()
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    labels, _, inertia = K.assign_nearest(X, C, "euclidean")
    e1.record()
    e1.synchronize()
    print("assign %d x %d -> k=%d: %.3f ms" % (n, D, k, e0.elapsed_time(e1)), flush=True)
