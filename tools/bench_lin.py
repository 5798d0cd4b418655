"""Micro-benchmark of lin_tc (pre-split weights) with bring-up flags (run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from framedipt_b200 import runtime
ctx = runtime.Context()
for name, M, N, K in [("proj", 2800, 6816, 256), ("256x256", 2800, 256, 256), ("in_proj", 2800, 960, 320), ("320x320", 2800, 320, 320),
                      ("K=128 N=384", 2800, 384, 128), ("cfg3 proj", 16384, 6816, 256)]:
    x = torch.randn(M, K, device="cuda"); w = torch.randn(N, K, device="cuda"); b = torch.randn(N, device="cuda")
    res = []
    for flags in (0, 1, 2, 3, 1024):
        ctx.set_option(3, flags)
        res.append(ctx.bench_linear(x, w, b))
    ctx.set_option(3, 0)
    print(f"{name:12s} M={M} N={N} K={K}: full {res[0]:7.1f} us | no stores {res[1]:7.1f} | no MMA {res[2]:7.1f} | neither {res[3]:7.1f} | stg16 {res[4]:7.1f}  "
          f"({2.0*M*N*K/res[0]/1e6:.1f} TFLOP/s)", flush=True)
