"""clock64 timeline of worker thread 0 of CTA 0 of the edge-embedder kernel on the cfg2 shape (run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from framedipt_b200 import SE3Diffuser, synthetic
from framedipt_b200.config import default_conf
from framedipt_b200.params import synthetic_state_dict
from framedipt_b200.score_network import ScoreNetwork
conf = default_conf()
diffuser = SE3Diffuser(conf.diffuser)
m = ScoreNetwork(conf.model, diffuser, inpainting=True)
m.load_state_dict(synthetic_state_dict(0))
m = m.to("cuda").eval()
wl = synthetic.WORKLOADS["cfg2_tcr350"]
np.random.seed(1)
feats = {k: v.to("cuda") for k, v in synthetic.make_features(wl, diffuser, seed=0).items()}
ctx = m.context(torch.device("cuda", 0))
pf = m.prepare(feats, torch.device("cuda", 0))
ctx.set_option(2, 1)
ctx.set_option(3, 8192)
for _ in range(2):
    ctx.embed(pf, feats["t"])
torch.cuda.synchronize()
ts = ctx.debug_read(8 * 16).reshape(8, 16)
ctx.set_option(3, 0)
names = ["tile start", "d0_full seen", "E0 math done", "A1 free (barrier)", "E0 stored", "next features built", "d1_full seen", "E1 stored",
         "d2_full seen", "E2 stats done", "E2 stored", "store barrier passed"]
for tile in (2, 3):
    t0 = int(ts[tile][0])
    prev = 0
    print(f"--- tile {tile}")
    for i, n in enumerate(names):
        t = int(ts[tile][i]) - t0
        print(f"{t:8d} (+{t - prev:5d})  {n}")
        prev = t
print("tile period (cycles):", [int(ts[i + 1][0] - ts[i][0]) for i in range(1, 7)])
