"""TEST INFRASTRUCTURE ONLY — generates tests/golden/logp_small.npz by running the UNMODIFIED reference's EigenFold confidence
score path (experiments/utils.py:752-869: SE3Diffuser.forward / log_prob_forward / log_prob_backward + one_step_inference_score).

Run in the build container (needs /root/reference):   python oracle/make_golden_logp.py
"""
from __future__ import annotations

import copy
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
    from framedipt_b200 import synthetic
    from framedipt_b200.params import synthetic_state_dict
    from framedipt_b200.synthetic import Workload

    conf = rh.default_conf(input_aatype=True, seed=123)
    sd = synthetic_state_dict(0)
    diffuser = se3.SE3Diffuser(conf.diffuser)
    model = sn.ScoreNetwork(conf.model, diffuser, inpainting=True)
    model.load_state_dict(sd, strict=True)
    model.eval()

    wl = Workload("logp24", 1, (14, 10), ((5, 10), (17, 20)), 10)
    np.random.seed(123)
    feats = synthetic.make_features(wl, diffuser, seed=4)
    save = {f"in_{k}": v.numpy() for k, v in feats.items()}
    diffuse_mask = ((1 - feats["fixed_mask"]) * feats["res_mask"])[0].numpy().astype(np.float64)
    rig0 = ru.Rigid.from_tensor_7(feats["rigids_t"][0].clone())

    # ---- diffuser-level known answers: one forward-noising step + both log-probabilities, for several (t_1, dt)
    cases = []
    rs = np.random.RandomState(17)
    for ci, (t_1, dt) in enumerate([(0.01, 0.1), (0.3, 0.1), (0.62, 0.02), (0.9, 0.1)]):
        np.random.seed(1000 + ci)
        rig_a = copy.deepcopy(rig0) if ci % 2 == 0 else cases[-1]["rig_b"]
        rig_b = diffuser.forward(rigids_t_1=copy.deepcopy(rig_a), t_1=t_1, diffuse_mask=diffuse_mask, dt=dt)
        t = min(t_1 + dt, 1.0)
        ts = rs.normal(size=(wl.n_res, 3)) * 3.0
        rsx = rs.normal(size=(wl.n_res, 3)) * 2.0
        lpf = diffuser.log_prob_forward(rigids_t=rig_b, rigids_t_1=rig_a, dt=dt, t_1=t_1, diffuse_mask=diffuse_mask)
        lpb = diffuser.log_prob_backward(rigids_t=rig_b, rigids_t_1=rig_a, trans_score_t=ts, rot_score_t=rsx, dt=dt, t=t,
                                         diffuse_mask=diffuse_mask)
        cases.append(dict(rig_b=rig_b))
        save[f"c{ci}_scalars"] = np.array([t_1, dt, t, float(lpf), float(lpb)], np.float64)
        save[f"c{ci}_a_rot"] = rig_a.get_rots().get_rot_mats().numpy()
        save[f"c{ci}_a_trans"] = rig_a.get_trans().numpy()
        save[f"c{ci}_b_rot"] = rig_b.get_rots().get_rot_mats().numpy()
        save[f"c{ci}_b_trans"] = rig_b.get_trans().numpy()
        save[f"c{ci}_trans_score"] = ts
        save[f"c{ci}_rot_score"] = rsx
        print(f"case {ci}: t_1={t_1} dt={dt} log q = {lpf:.6f}  log p = {lpb:.6f}  dtypes rot {save[f'c{ci}_b_rot'].dtype} trans {save[f'c{ci}_b_trans'].dtype}")
    save["n_cases"] = np.array(len(cases))

    # ---- the whole confidence score with the synthetic network
    for tag, num_t, min_t in (("s6", 6, 0.01), ("s12", 12, 0.05)):
        np.random.seed(77)
        f2 = {k: v.clone() for k, v in feats.items()}
        lp, lps = eu.logp_confidence_score(model=model, diffuser=diffuser, rigids_t=copy.deepcopy(rig0), sample_feats=f2,
                                           diffuse_mask=diffuse_mask, num_t=num_t, min_t=min_t, device="cpu", self_condition=True)
        save[f"{tag}_args"] = np.array([num_t, min_t], np.float64)
        save[f"{tag}_log_prob"] = np.array(float(lp), np.float64)
        save[f"{tag}_log_probs"] = np.array([float(x) for x in lps], np.float64)
        print(tag, "log_prob", float(lp), "per step", [round(float(x), 3) for x in lps])
    np.savez_compressed(os.path.join(OUT, "logp_small.npz"), **save)


if __name__ == "__main__":
    main()
