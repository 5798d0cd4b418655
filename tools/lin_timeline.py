"""clock64 timeline of CTA (0,0) of one lin_tc launch for typical node-side layers (run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from framedipt_b200 import runtime
ctx = runtime.Context()
ctx.set_option(2, 1)
names = ["kernel start", "setup done (barriers, TMEM)", "W pdl_wait done", "W first X loads issued", "W a_full[0] arrived", "W all X staged",
         "M a_full[0] seen", "M first weights seen", "M tile 0 MMAs issued", "W acc_full seen", "W TMEM loaded", "W stores issued", "all warps done"]
for name, M, N, K in [("256x256", 2800, 256, 256), ("320x320", 2800, 320, 320), ("in_proj 960", 2800, 960, 320), ("N=128 K=256", 2800, 128, 256)]:
    x = torch.randn(M, K, device="cuda"); w = torch.randn(N, K, device="cuda"); b = torch.randn(N, device="cuda")
    ctx.set_option(3, int(os.environ.get("LT_FLAG", "512")))
    us = ctx.bench_linear(x, w, b)
    ts = ctx.debug_read(16)
    ctx.set_option(3, 0)
    us0 = ctx.bench_linear(x, w, b)
    print(f"--- {name}: {us0:.1f} us per launch back-to-back ({us:.1f} with stamps)")
    ev = sorted((int(ts[i] - ts[0]), names[i]) for i in range(13) if ts[i])
    prev = 0
    for t, n in ev:
        print(f"{t:8d} (+{t - prev:6d})  {n}")
        prev = t
