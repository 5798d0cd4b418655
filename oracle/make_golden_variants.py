"""TEST INFRASTRUCTURE ONLY — tests/golden/traj_variants.npz: short trajectories of the UNMODIFIED reference with the sampler's option
switches away from their defaults (center, noise_scale, self_condition, diffuse_rot / diffuse_trans); they pin the matching switches of
the oracle restatement (tests/test_oracle_golden.py) and, through it, of the CUDA path.

Run in the build container (needs /root/reference):   python oracle/make_golden_variants.py
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
CASES = {  # tag: (center, noise_scale, self_condition, diffuse_rot, diffuse_trans)
    "a": (False, 1.0, False, True, True),
    "b": (True, 0.3, True, False, True),
    "c": (True, 0.3, True, True, False),
}


def main():
    torch.set_num_threads(8)
    sn, se3, eu, ru, aa = rh.load_reference()
    from framedipt_b200 import synthetic
    from framedipt_b200.params import synthetic_state_dict
    from framedipt_b200.synthetic import Workload

    sd = synthetic_state_dict(0)
    wl = Workload("var24", 2, (14, 10), ((5, 10), (17, 20)), 8)
    save = {}
    for tag, (center, ns, sc, drot, dtrans) in CASES.items():
        conf = rh.default_conf(input_aatype=True, seed=123)
        conf.diffuser.diffuse_rot, conf.diffuser.diffuse_trans = drot, dtrans
        diffuser = se3.SE3Diffuser(conf.diffuser)
        model = sn.ScoreNetwork(conf.model, diffuser, inpainting=True)
        model.load_state_dict(sd, strict=True)
        model.eval()
        np.random.seed(123)
        feats = synthetic.make_features(wl, se3.SE3Diffuser(rh.default_conf(input_aatype=True, seed=123).diffuser), seed=6)
        np.random.seed(4242)
        st = np.random.get_state()
        n_draw = (1 if drot else 0) + (1 if dtrans else 0)
        raw = np.random.normal(size=(wl.num_t - 1, n_draw, wl.batch, wl.n_res, 3))
        noise = np.zeros((wl.num_t - 1, 2, wl.batch, wl.n_res, 3))
        k = 0
        if drot:
            noise[:, 0] = raw[:, k]
            k += 1
        if dtrans:
            noise[:, 1] = raw[:, k]
        np.random.set_state(st)
        out = eu.inference_fn(model, diffuser, feats, num_t=wl.num_t, min_t=0.01, center=center, aux_traj=True, self_condition=sc,
                              noise_scale=ns, inpainting=True, input_aatype=True)
        if tag == "a":
            save.update({f"in_{k2}": v.numpy() for k2, v in feats.items()})
        save[f"{tag}_opts"] = np.array([center, ns, sc, drot, dtrans], np.float64)
        save[f"{tag}_noise"] = noise
        save[f"{tag}_prot_traj"] = out["prot_traj"][:, :, :, :5].astype(np.float32)
        save[f"{tag}_rigid_traj"] = out["rigid_traj"]
        print(tag, CASES[tag], "final CA[0,0]", out["prot_traj"][0, 0, 0, 1])
    save["num_t"] = np.array(wl.num_t)
    np.savez_compressed(os.path.join(OUT, "traj_variants.npz"), **save)


if __name__ == "__main__":
    main()
