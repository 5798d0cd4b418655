"""Prints the clock64 timeline of CTA 0 of one EdgeTransition call on the cfg2 shape (run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from framedipt_b200 import runtime
from framedipt_b200.params import synthetic_state_dict
ctx = runtime.Context()
ctx.load_state_dict(synthetic_state_dict(0))
ctx.set_option(2, 1)
ctx.set_option(10, int(os.environ.get("R2_TMEM", "1")))
ctx.set_option(3, int(os.environ.get("DBG_FLAGS", "0")))
B, N = 8, 350
node = torch.randn(B, N, 256, device="cuda"); z = torch.randn(B, N, N, 128, device="cuda"); mask = torch.ones(B, N, device="cuda")
for _ in range(2):
    ctx.edge_transition(0, node, z, mask)
ts = ctx.debug_read().reshape(8, 48)
names = {0: "M tile start", 42: "M G1(0) kb0 local landed", 43: "M G1(0) kb0 peer landed", 44: "M G1(0) kb1 local landed", 45: "M G1(0) kb1 peer landed", 46: "M G1(0) kb2 local landed", 47: "M G1(0) kb2 peer landed", 33: "W E3 D3 loaded", 34: "W E3 partial stats done", 35: "W E3 barrier passed", 36: "L G1(0) kb3 TMA issued", 37: "M G1(0) kb1 issued", 38: "M G1(0) kb2 issued", 39: "M G1(0) kb3 issued", 14: "M G1(0) ds_empty ok", 15: "M G1(0) kb0 issued", 1: "M G1(0) issued", 2: "M G1(1) issued", 3: "M G2(0) issued", 4: "M G1(2) issued", 5: "M G2(1) issued", 6: "M G2(2) issued",
         7: "M G3s issued", 11: "M G3p0 bufwait done", 8: "M G3p(0) issued", 9: "M G3p(1) issued", 10: "M G3p(2) issued",
         16: "W tile start", 17: "W E1(0) ds_full", 18: "W E1(0) math done", 19: "W E1(0) stored", 20: "W E1(1) ds_full", 21: "W E1(1) math done",
         22: "W E1(1) stored", 23: "W E1(2) ds_full", 24: "W E1(2) math done", 25: "W E1(2) stored", 26: "W E2 d2_full", 27: "W E2(0) stored",
         28: "W E2(1) stored", 29: "W E2(2) stored", 30: "W E3 ds_full", 31: "W E3 LN done", 32: "W E3 store done"}
t0 = ts[1][0]
for tile in (() if os.environ.get('SHORT') else (1, 2)):
    ev = sorted((int(ts[tile][k]) - int(t0), names[k]) for k in names if ts[tile][k] != 0)
    print(f"--- tile {tile}")
    prev = None
    for t, n in ev:
        print(f"{t:8d} (+{0 if prev is None else t - prev:5d})  {n}")
        prev = t
print("weight-stage wait cycles (cumulative) local/peer/MMA-issue/commit:", [(int(ts[i][40]), int(ts[i][41]), int(ts[i][13]), int(ts[i][12])) for i in range(0, 6)])
print("tile period (cycles):", [int(ts[i + 1][0] - ts[i][0]) for i in range(1, 6)])
