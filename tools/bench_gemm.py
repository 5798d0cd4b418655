"""Micro-benchmark of the node-side GEMM kernel on the shapes of cfg2 (run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from framedipt_b200 import runtime

ctx = runtime.Context()
shapes = [("node 256x256", 1, 2800, 256, 256, True), ("tfmr in_proj", 1, 2800, 960, 320, True), ("tfmr 320x320", 1, 2800, 320, 320, True),
          ("ipa proj", 1, 2800, 6816, 256, True), ("ipa linear_out", 1, 2800, 256, 2688, True), ("ipa S", 64, 350, 350, 280, True),
          ("ipa AV", 64, 350, 292, 350, False), ("tfmr S", 32, 350, 350, 80, True), ("tfmr PV", 32, 350, 80, 350, False),
          ("edge chunk", 1, 1 << 20, 128, 128, True), ("cfg3 proj", 1, 16384, 6816, 256, True)]
for use_tc in (1, 2):
    ctx.set_option(3, 8 if use_tc == 2 else 0)
    print("== gemm_tc bn=64 (2 CTAs/SM)" if use_tc == 1 else "== gemm_tc bn=128 where N > 128")
    for name, Bt, M, N, K, km in shapes:
        a = torch.randn(Bt, M, K, device="cuda")
        b = torch.randn(Bt, N, K, device="cuda") if km else torch.randn(Bt, K, N, device="cuda")
        for _ in range(3):
            ctx.matmul(a, b, km)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20
        e0.record()
        for _ in range(reps):
            ctx.matmul(a, b, km)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / reps
        print(f"  {name:16s} B={Bt:3d} M={M:7d} N={N:5d} K={K:5d}: {us:8.1f} us  {2.0*Bt*M*N*K/us/1e6:7.1f} TFLOP/s", flush=True)
