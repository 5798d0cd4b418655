"""TEST INFRASTRUCTURE ONLY — tests/golden/config_variants.npz: the UNMODIFIED reference with the configuration switches that
change the arithmetic of the hot path away from their defaults (SURVEY.md §8b "Variant switches"), plus a padded mixed-length batch:

  cached   diffuser.so3.use_cached_score=True (so3_diffuser.py:389-396): rot-score grid, one forward, a 6-step trajectory
  nosc     model.embed.embed_self_conditioning=False (score_network.py:95-96, 185; experiments/utils.py:356-358, 571-578):
           edge embedder without the distogram; forward + 6-step trajectory through inference_fn(embed_self_conditioning=False)
  bbflags  inference_fn(inpainting=False, input_aatype=False) on an aatype-taking model (experiments/utils.py:549-555): the
           trajectory's backbone atoms use ALA frames everywhere (GLY gets a CB), 4 steps
  mixed    two DIFFERENT structures (N=20 and N=31) padded to 31 with the reference's own pad_feats / pad_rigid
           (framedipt/data/utils.py:311-339) and batched: forward + 6-step trajectory

Run in the build container (needs /root/reference):   python oracle/make_golden_config.py
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_harness as rh  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def main():
    torch.set_num_threads(8)
    sn, se3, eu, ru, aa = rh.load_reference()
    from framedipt.data import utils as du  # type: ignore  (the reference's)
    from framedipt_b200 import synthetic
    from framedipt_b200.params import ModelDims, synthetic_state_dict
    from framedipt_b200.synthetic import Workload

    save = {}
    wl = Workload("cv24", 2, (14, 10), ((5, 10), (17, 20)), 6)

    def traj(model, diffuser, feats, num_t, **kw):
        st = np.random.get_state()
        noise = synthetic.draw_noise(num_t, feats["res_mask"].shape[0], feats["res_mask"].shape[1])
        np.random.set_state(st)
        out = eu.inference_fn(model, diffuser, feats, num_t=num_t, min_t=0.01, aux_traj=True, noise_scale=0.1, **kw)
        return noise, out

    # ------------------------------------------------------------------ cached score table
    conf = rh.default_conf(input_aatype=True, seed=123)
    conf.diffuser.so3.use_cached_score = True
    diffuser = se3.SE3Diffuser(conf.diffuser)
    sd = synthetic_state_dict(0)
    model = sn.ScoreNetwork(conf.model, diffuser, inpainting=True)
    model.load_state_dict(sd, strict=True)
    model.eval()
    grid = dict(np.load(os.path.join(OUT, "scores_grid.npz")))
    q_t, q_0r, tt = torch.tensor(grid["q_t"]), torch.tensor(grid["q_0"]), torch.tensor(grid["t"])
    q_id = torch.zeros_like(q_t)
    q_id[..., 0] = 1.0
    with torch.no_grad():
        s_id = diffuser.calc_rot_score(ru.Rotation(quats=q_t, normalize_quats=False), ru.Rotation(quats=q_id, normalize_quats=False), tt)
        s_rd = diffuser.calc_rot_score(ru.Rotation(quats=q_t, normalize_quats=False), ru.Rotation(quats=q_0r, normalize_quats=False), tt)
    save.update(cached_rot_score_identity0=s_id.numpy(), cached_rot_score_random0=s_rd.numpy())
    # a few rows of the table itself pin the host-side table builder
    so3 = diffuser._so3_diffuser
    save.update(cached_table_rows=np.array([0, 1, 250, 500, 999]), cached_table=so3._score_norms[[0, 1, 250, 500, 999]])
    np.random.seed(123)
    feats = synthetic.make_features(wl, diffuser, seed=3)
    f1 = {k: v.clone() for k, v in feats.items()}
    f1["t"] = torch.tensor([0.37, 0.81])
    with torch.no_grad():
        o = model(f1)
    save.update({f"cached_in_{k}": v.numpy() for k, v in feats.items()})
    save.update(cached_fwd_rot_score=o["rot_score"].numpy(), cached_fwd_rigids=o["rigids"].numpy())
    noise, out = traj(model, diffuser, feats, wl.num_t, inpainting=True, input_aatype=True)
    save.update(cached_noise=noise, cached_prot_traj=out["prot_traj"][:, :, :, :5].astype(np.float32))
    print("cached ok", s_id.dtype, flush=True)

    # ------------------------------------------------------------------ no self-conditioning features
    conf = rh.default_conf(input_aatype=True, seed=123)
    conf.model.embed.embed_self_conditioning = False
    diffuser = se3.SE3Diffuser(conf.diffuser)
    sd_nosc = synthetic_state_dict(0, ModelDims(embed_self_conditioning=False))
    model = sn.ScoreNetwork(conf.model, diffuser, inpainting=True)
    model.load_state_dict(sd_nosc, strict=True)
    model.eval()
    np.random.seed(123)
    feats = synthetic.make_features(wl, diffuser, seed=3)
    f1 = {k: v.clone() for k, v in feats.items()}
    f1["t"] = torch.tensor([0.37, 0.81])
    f1["sc_ca_t"] = torch.randn(2, wl.n_res, 3, generator=torch.Generator().manual_seed(1)) * 5  # must have no effect
    with torch.no_grad():
        o = model(f1)
    save.update({f"nosc_in_{k}": v.numpy() for k, v in feats.items()})
    save.update(nosc_fwd_sc_ca_t=f1["sc_ca_t"].numpy(), nosc_fwd_rigids=o["rigids"].numpy(), nosc_fwd_trans_score=o["trans_score"].numpy(),
                nosc_fwd_psi=o["psi"].numpy())
    noise, out = traj(model, diffuser, feats, wl.num_t, inpainting=True, input_aatype=True, embed_self_conditioning=False)
    save.update(nosc_noise=noise, nosc_prot_traj=out["prot_traj"][:, :, :, :5].astype(np.float32))
    print("nosc ok", flush=True)

    # ------------------------------------------------------------------ backbone residue types follow the call's flags
    conf = rh.default_conf(input_aatype=True, seed=123)
    diffuser = se3.SE3Diffuser(conf.diffuser)
    model = sn.ScoreNetwork(conf.model, diffuser, inpainting=True)
    model.load_state_dict(sd, strict=True)
    model.eval()
    np.random.seed(123)
    feats = synthetic.make_features(wl, diffuser, seed=3)
    feats["aatype"][0, :4] = torch.tensor([7, 7, 14, 0])  # GLY, GLY, PRO, ALA
    for tag, (inp, ia) in {"bbff": (False, False), "bbtf": (True, False)}.items():
        np.random.seed(55)
        noise, out = traj(model, diffuser, feats, 4, inpainting=inp, input_aatype=ia)
        save.update({f"{tag}_noise": noise, f"{tag}_prot_traj": out["prot_traj"][:, :, :, :5].astype(np.float32),
                     f"{tag}_rigid_0_traj": out["rigid_0_traj"][:, :, :, :5].astype(np.float32)})
    save.update({f"bb_in_{k}": v.numpy() for k, v in feats.items()})
    print("bbflags ok", flush=True)

    # ------------------------------------------------------------------ padded mixed-length batch
    wa = Workload("mixA", 1, (12, 8), ((3, 8),), 6)
    wb = Workload("mixB", 1, (31,), ((10, 19),), 6)
    np.random.seed(123)
    fa = synthetic.make_features(wa, diffuser, seed=41)
    fb = synthetic.make_features(wb, diffuser, seed=42)
    pa = du.pad_feats({k: v[0] for k, v in fa.items()}, 31, use_torch=True)
    pb = du.pad_feats({k: v[0] for k, v in fb.items()}, 31, use_torch=True)
    feats = {k: torch.stack([pa[k], pb[k]]) for k in pa if k != "t"}
    feats["t"] = torch.ones(2)
    f1 = {k: v.clone() for k, v in feats.items()}
    f1["t"] = torch.tensor([0.6, 0.6])
    f1["sc_ca_t"] = f1["rigids_t"][..., 4:].float() * f1["res_mask"][..., None].float()
    with torch.no_grad():
        o = model(f1)
    save.update({f"mixed_in_{k}": v.numpy() for k, v in feats.items()})
    save.update(mixed_fwd_sc_ca_t=f1["sc_ca_t"].numpy(), mixed_fwd_rigids=o["rigids"].numpy(), mixed_fwd_trans_score=o["trans_score"].numpy(),
                mixed_fwd_rot_score=o["rot_score"].numpy(), mixed_fwd_psi=o["psi"].numpy())
    np.random.seed(77)
    noise, out = traj(model, diffuser, feats, 6, inpainting=True, input_aatype=True)
    save.update(mixed_noise=noise, mixed_prot_traj=out["prot_traj"][:, :, :, :5].astype(np.float32))
    # each structure alone (unpadded): what the padded batch must reproduce on its valid residues
    for tag, f, n in (("mixA", fa, 0), ("mixB", fb, 1)):
        with torch.no_grad():
            g1 = {k: v.clone() for k, v in f.items()}
            g1["t"] = torch.tensor([0.6])
            g1["sc_ca_t"] = g1["rigids_t"][..., 4:].float()
            oo = model(g1)
        save[f"{tag}_alone_fwd_rigids"] = oo["rigids"].numpy()
    print("mixed ok", flush=True)
    np.savez_compressed(os.path.join(OUT, "config_variants.npz"), **save)


if __name__ == "__main__":
    main()
