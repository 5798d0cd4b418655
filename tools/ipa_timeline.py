"""Prints the clock64 timeline of CTA 0 of one ipa_core launch on the cfg2 shape (run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from framedipt_b200 import runtime
from framedipt_b200.params import synthetic_state_dict
ctx = runtime.Context()
ctx.load_state_dict(synthetic_state_dict(0))
ctx.set_option(2, 1)   # allocate the timeline buffer
ctx.set_option(3, 4)   # route it to ipa_core
B, N = 8, 350
s = torch.randn(B, N, 256, device="cuda"); z = torch.randn(B, N, N, 128, device="cuda"); mask = torch.ones(B, N, device="cuda")
q = torch.randn(B, N, 4, device="cuda"); q = q / q.norm(dim=-1, keepdim=True); tr = torch.randn(B, N, 3, device="cuda")
for _ in range(2):
    ctx.ipa(0, s, z, q, tr, mask)
ts = ctx.debug_read().reshape(8, 48)
names = {1: "M GEMM-b t0 issued", 9: "M p_full seen", 10: "M GEMM-o issued", 20: "E row start (softmax)", 31: "E softmax done, P stored",
         30: "E logits(next) done", 32: "E d2_full seen", 33: "E down_z done"}
t0 = ts[1][20]
for it in (1, 2, 3):
    ev = sorted((int(ts[it][k]) - int(t0), names[k]) for k in names if ts[it][k] != 0)
    print(f"--- row iteration {it}")
    prev = None
    for t, n in ev:
        print(f"{t:8d} (+{0 if prev is None else t - prev:5d})  {n}")
        prev = t
print("row period (cycles, epilogue):", [int(ts[i + 1][20] - ts[i][20]) for i in range(1, 6)])
