"""A/B of the EdgeTransition kernel: CTA-pair (et_fused2.cuh) vs single-CTA (et_fused.cuh) -- equality of results and timing (GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from framedipt_b200 import runtime
from framedipt_b200.params import synthetic_state_dict
ctx = runtime.Context()
ctx.load_state_dict(synthetic_state_dict(0))
OPT_ET_PAIR = 5
if not os.environ.get("NO_TIMELINE"):
    ctx.set_option(2, 1)  # host-mapped timeline + barrier-timeout records
BAR_NAMES = ["w_full"] * 7 + ["w_peer"] * 7 + ["w_empty"] * 7 + ["az_full"] * 2 + ["az_peer"] * 2 + ["az_empty"] * 2 + ["an_full", "an_peer", "ds_full", "ds_empty"] + \
    ["buf_full"] * 2 + ["buf_free"] * 2 + ["d2_full", "d2_empty"] + ["vec_full"] * 2 + ["vec_free"] * 2 + ["stg_full"]


def dump_failures():
    import numpy as np
    d = ctx.debug_read(8 * 48 + 32)
    fb = d[8 * 48:].astype(np.uint64)
    print("barrier timeouts recorded:", int(fb[0]))
    offs = sorted(set(int((int(x) >> 1) & 0x3FFFFF) for x in fb[1:1 + min(int(fb[0]), 31)]))
    base = None
    for x in fb[1:1 + min(int(fb[0]), 31)]:
        x = int(x)
        print(f"  block {x >> 48} thread {(x >> 32) & 0xFFFF} (warp {((x >> 32) & 0xFFFF) // 32}) bar smem offset {(x >> 1) & 0x3FFFFF:#x} cluster-wait {(x >> 24) & 1} parity {x & 1}")
    print("timeline tile 0:", [int(v) for v in d[:48]])


def run(pair, node, z, mask, reps=0):
    ctx.set_option(OPT_ET_PAIR, pair)
    out = ctx.edge_transition(0, node, z, mask)
    torch.cuda.synchronize()
    ms = None
    if reps:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            ctx.edge_transition(0, node, z, mask)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
    return out, ms


g = torch.Generator(device="cuda").manual_seed(0)
for (B, N, reps) in [(1, 37, 0), (1, 131, 0), (2, 129, 0), (3, 255, 0), (8, 350, 20), (64, 256, 5)]:
    node = torch.randn(B, N, 256, device="cuda", generator=g)
    z = torch.randn(B, N, N, 128, device="cuda", generator=g)
    mask = (torch.rand(B, N, device="cuda", generator=g) > 0.1).float()
    if os.environ.get("PAIR_FIRST"):
        try:
            run(1, node, z, mask, 0)
        except Exception as e:
            print("pair kernel failed:", str(e).splitlines()[0])
            dump_failures()
            raise SystemExit(1)
    o1, t1 = run(0, node, z, mask, reps)
    try:
        o2, t2 = run(1, node, z, mask, reps)
    except Exception as e:
        print("pair kernel failed:", str(e).splitlines()[0])
        dump_failures()
        raise SystemExit(1)
    d = (o1 - o2).abs().max().item()
    print(f"B={B} N={N}: max|pair - single| = {d:.3e}  single {t1} ms  pair {t2} ms (includes the per-residue GEMMs + fp32<->image conversions)", flush=True)
    assert d == 0.0 or d < 1e-6, d
