"""Bring-up probe for gemm_tc (run on the GPU box): python tools/bringup_gemm_tc.py <case>"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from framedipt_b200 import runtime

case = sys.argv[1]
ctx = runtime.Context()
g = torch.Generator().manual_seed(0)


def check(Bt, M, N, K, kmajor, pad=0):
    a = torch.randn(Bt, M, K + pad, generator=g).cuda()[:, :, :K]
    b = (torch.randn(Bt, N, K + pad, generator=g).cuda()[:, :, :K] if kmajor else torch.randn(Bt, K, N + pad, generator=g).cuda()[:, :, :N])
    c = ctx.matmul(a, b, kmajor)
    ref = (a.double() @ (b.double().transpose(1, 2) if kmajor else b.double()))
    err = (c.double() - ref).abs().max().item()
    scale = ref.abs().max().item()
    print(f"  B={Bt} M={M} N={N} K={K} kmajor={kmajor} pad={pad}: max err {err:.3e} (scale {scale:.2f}) rel {err/scale:.2e}", flush=True)
    return err / scale


if case == "kmajor":
    r = [check(1, 128, 128, 32, True), check(1, 300, 256, 256, True), check(3, 350, 350, 280, True), check(2, 257, 70, 86, True, pad=1),
         check(1, 2800, 6816, 256, True)]
elif case in ("mn0", "mn1"):
    ctx.set_option(1, int(case[-1]))
    r = [check(1, 128, 128, 32, False), check(1, 128, 64, 8, False), check(2, 350, 292, 350, False), check(1, 300, 80, 301, False, pad=3)]
elif case == "simt":
    ctx.set_option(0, 0)
    r = [check(2, 350, 292, 350, False), check(3, 350, 350, 280, True)]
torch.cuda.synchronize()
print(case, "WORST", max(r), flush=True)
