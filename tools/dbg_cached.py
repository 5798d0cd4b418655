import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from framedipt_b200 import SE3Diffuser, Rotation
from framedipt_b200.config import default_conf
from oracle import framedipt_oracle as orc
G = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
g = np.load(os.path.join(G, "config_variants.npz")); grid = np.load(os.path.join(G, "scores_grid.npz"))
conf = default_conf(); conf.diffuser.so3.use_cached_score = True; conf.diffuser.so3.cache_dir = "/tmp/igso3_dbg"
d = SE3Diffuser(conf.diffuser)
q_t, tt = torch.tensor(grid["q_t"]), torch.tensor(grid["t"])
q0 = torch.zeros_like(q_t); q0[..., 0] = 1
s = d.calc_rot_score(Rotation(quats=q_t), Rotation(quats=q0), tt).cpu().numpy()
ref = g["cached_rot_score_identity0"]
v = orc.quat_to_rotvec(orc.quat_multiply(q0 * torch.tensor([1., -1, -1, -1]), q_t)); om = torch.linalg.norm(v, dim=-1).numpy()
scale = np.abs(ref).max(-1, keepdims=True) + 1e-30
bad = (np.abs(s - ref) / scale > 1e-5).any(-1)
print("bad per t row", bad.sum(-1), "of", bad.shape)
for (i, j) in list(zip(*np.nonzero(bad)))[:12]:
    print(i, j, "omega", om[i, j], "ours", s[i, j], "ref", ref[i, j], "ratio", s[i, j] / ref[i, j])
tab = d._so3_diffuser.score_norms
print("table rows vs ref", [float(np.abs(tab[r] - row).max()) for r, row in zip(g["cached_table_rows"], g["cached_table"])])
