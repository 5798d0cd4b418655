"""TEST INFRASTRUCTURE ONLY — generates tests/golden/*.npz by running the UNMODIFIED reference.

Run in the build container (needs /root/reference):   python oracle/make_golden.py
The fixtures pin both the oracle restatement (tests/test_oracle_golden.py, CPU) and the CUDA path
(tests/test_gpu_*.py).  Inputs are regenerated deterministically from framedipt_b200.synthetic /
framedipt_b200.params, and are also stored so the fixtures are self-contained.
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_harness as rh  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def small_workload():
    from framedipt_b200.synthetic import Workload

    return Workload("small24", 2, (14, 10), ((5, 10), (17, 20)), 10)


def build_model(sn, se3, conf, sd):
    diffuser = se3.SE3Diffuser(conf.diffuser)
    model = sn.ScoreNetwork(conf.model, diffuser, inpainting=True)
    missing = model.load_state_dict(sd, strict=True)
    model.eval()
    return model, diffuser


def main():
    torch.set_num_threads(8)
    os.makedirs(OUT, exist_ok=True)
    sn, se3, eu, ru, aa = rh.load_reference()
    from framedipt_b200 import synthetic
    from framedipt_b200.params import synthetic_state_dict

    conf = rh.default_conf(input_aatype=True, seed=123)
    sd = synthetic_state_dict(0)
    t0 = time.time()
    model, diffuser = build_model(sn, se3, conf, sd)
    print(f"reference model + diffuser ready in {time.time() - t0:.1f}s")

    # ------------------------------------------------------------------ A: single forward, with taps
    for tag, pad in (("forward_small", 0), ("forward_small_padded", 3)):
        wl = small_workload()
        np.random.seed(123)
        feats = synthetic.make_features(wl, diffuser, seed=3)
        rs = np.random.RandomState(7)
        feats["sc_ca_t"] = torch.tensor(rs.normal(size=(wl.batch, wl.n_res, 3)) * 6.0, dtype=torch.float32)
        feats["sc_ca_t"][0, 3] = feats["sc_ca_t"][0, 4]  # a zero-distance off-diagonal pair
        feats["t"] = torch.tensor([0.37, 0.81], dtype=torch.float32)
        if pad:
            feats["res_mask"][1, -pad:] = 0.0
        taps = {}
        hooks = []

        def tap(name, sel=lambda o: o):
            def fn(_m, _i, o):
                taps[name] = sel(o).detach().clone().numpy()
            return fn

        hooks.append(model.embedding_layer.register_forward_hook(lambda m, i, o: taps.update(node_embed_raw=o[0].detach().numpy().copy(), edge_embed_raw=o[1].detach().numpy().copy())))
        tr = model.score_model.trunk
        for b in range(4):
            hooks.append(tr[f"ipa_{b}"].register_forward_hook(tap(f"ipa_{b}")))
            hooks.append(tr[f"node_transition_{b}"].register_forward_hook(tap(f"node_transition_{b}")))
            hooks.append(tr[f"seq_tfmr_{b}"].register_forward_hook(tap(f"seq_tfmr_{b}")))
            if b < 3:
                hooks.append(tr[f"edge_transition_{b}"].register_forward_hook(tap(f"edge_transition_{b}")))
        with torch.no_grad():
            out = model({k: v.clone() for k, v in feats.items()})
        for h in hooks:
            h.remove()
        save = {f"in_{k}": v.numpy() for k, v in feats.items()}
        save.update({f"out_{k}": out[k].detach().numpy() for k in ("rigids", "rot_score", "trans_score", "psi")})
        save["out_atom37"] = out["atom37"].detach().numpy()[:, :, :5]
        keep = ["node_embed_raw", "edge_embed_raw", "ipa_0", "ipa_3", "seq_tfmr_0", "node_transition_0", "node_transition_3", "edge_transition_0", "edge_transition_2"]
        save.update({f"tap_{k}": taps[k] for k in keep})
        np.savez_compressed(os.path.join(OUT, tag + ".npz"), **save)
        print(tag, {k: v.shape for k, v in save.items() if k.startswith("out_")})

    # ------------------------------------------------------------------ B: reverse step + backbone
    wl = small_workload()
    np.random.seed(123)
    feats = synthetic.make_features(wl, diffuser, seed=3)
    rs = np.random.RandomState(11)
    B, N = wl.batch, wl.n_res
    rot_score = rs.normal(size=(B, N, 3)) * 0.8
    trans_score = (rs.normal(size=(B, N, 3)) * 0.5).astype(np.float32)
    dmask = ((1 - feats["fixed_mask"]) * feats["res_mask"]).numpy()
    cases = {}
    for ci, (t, dt, ns, center) in enumerate([(0.63, 0.02, 0.1, True), (1.0, 0.01, 1.0, True), (0.05, 0.002, 0.5, False)]):
        st = np.random.get_state()
        z_rot = np.random.normal(size=(B, N, 3))
        z_trans = np.random.normal(size=(B, N, 3))
        np.random.set_state(st)
        rig = diffuser.reverse(rigid_t=ru.Rigid.from_tensor_7(feats["rigids_t"]), rot_score=rot_score, trans_score=trans_score,
                               diffuse_mask=dmask, t=t, dt=dt, center=center, noise_scale=ns)
        cases[f"c{ci}_params"] = np.array([t, dt, ns, float(center)])
        cases[f"c{ci}_z_rot"], cases[f"c{ci}_z_trans"] = z_rot, z_trans
        cases[f"c{ci}_rotmats"] = rig.get_rots().get_rot_mats().numpy()
        cases[f"c{ci}_trans"] = rig.get_trans().numpy()
        cases[f"c{ci}_tensor7"] = rig.to_tensor_7().numpy()
    psi = rs.normal(size=(B, N, 2)).astype(np.float32)
    psi /= np.linalg.norm(psi, axis=-1, keepdims=True)
    aat = feats["aatype"].clone()
    aat[0, :4] = torch.tensor([7, 20, 14, 0])  # GLY (no CB), unknown, PRO, ALA
    rig_in = ru.Rigid.from_tensor_7(feats["rigids_t"])
    atom37 = eu.get_atom_positions_from_rigids(rig_in, torch.tensor(psi), aat)
    np.savez_compressed(os.path.join(OUT, "reverse_small.npz"), rigids_t=feats["rigids_t"].numpy(), rot_score=rot_score,
                        trans_score=trans_score, diffuse_mask=dmask, psi=psi, aatype=aat.numpy(), atom37=atom37, **cases)
    print("reverse_small ok")

    # ------------------------------------------------------------------ C: rot-score / trans-score grid
    rs = np.random.RandomState(5)
    ax = rs.normal(size=(6, 40, 3))
    ax /= np.linalg.norm(ax, axis=-1, keepdims=True)
    ang = np.concatenate([np.array([1e-5, 5e-4, 1e-3, 2e-3, 0.01, 3.1, 3.14159]), rs.uniform(0.01, np.pi, 33)])
    hq = np.concatenate([np.cos(ang / 2)[None, :, None].repeat(6, 0), np.sin(ang / 2)[None, :, None] * ax], -1).astype(np.float32)
    q_t = torch.tensor(hq)
    q_0 = torch.zeros_like(q_t)
    q_0[..., 0] = 1.0
    q0r = rs.normal(size=(6, 40, 4)).astype(np.float32)
    q0r /= np.linalg.norm(q0r, axis=-1, keepdims=True)
    tt = torch.tensor([0.01, 0.05, 0.37, 0.63, 0.81, 1.0], dtype=torch.float32)
    with torch.no_grad():
        s_id = diffuser.calc_rot_score(ru.Rotation(quats=q_t, normalize_quats=False), ru.Rotation(quats=q_0, normalize_quats=False), tt)
        s_rand = diffuser.calc_rot_score(ru.Rotation(quats=q_t, normalize_quats=False), ru.Rotation(quats=torch.tensor(q0r), normalize_quats=False), tt)
        xt = torch.tensor(rs.normal(size=(6, 40, 3)).astype(np.float32) * 10)
        x0 = torch.tensor(rs.normal(size=(6, 40, 3)).astype(np.float32) * 10)
        ts = diffuser.calc_trans_score(xt, x0, tt[:, None, None], use_torch=True)
    np.savez_compressed(os.path.join(OUT, "scores_grid.npz"), q_t=hq, q_0=q0r, t=tt.numpy(), rot_score_identity0=s_id.numpy(),
                        rot_score_random0=s_rand.numpy(), trans_t=xt.numpy(), trans_0=x0.numpy(), trans_score=ts.numpy())
    print("scores_grid ok", s_id.dtype)

    # ------------------------------------------------------------------ D: sample_ref
    wl = synthetic.WORKLOADS["cfg1_monomer64"]
    np.random.seed(123)
    f_inp = synthetic.make_features(wl, diffuser, seed=0)
    wl3 = synthetic.Workload("denovo32", 2, (32,), (), 10, de_novo=True)
    with torch.no_grad():  # the reference's Rigid.identity() defaults to requires_grad=True
        f_dn = synthetic.make_features(wl3, diffuser, seed=0)
    np.savez_compressed(os.path.join(OUT, "sample_ref.npz"), inpaint_rigids_t=f_inp["rigids_t"].numpy(), denovo_rigids_t=f_dn["rigids_t"].numpy())
    print("sample_ref ok")

    # ------------------------------------------------------------------ E: config #1 trajectory (N=64, 50 steps)
    for tag, wl, seed in (("traj_cfg1", synthetic.WORKLOADS["cfg1_monomer64"], 0), ("traj_small", small_workload(), 3)):
        def run(nthreads):
            torch.set_num_threads(nthreads)
            np.random.seed(123)
            feats = synthetic.make_features(wl, diffuser, seed=seed)
            st = np.random.get_state()
            noise = synthetic.draw_noise(wl.num_t, wl.batch, wl.n_res)
            np.random.set_state(st)
            t0 = time.time()
            out = eu.inference_fn(model, diffuser, feats, num_t=wl.num_t, min_t=wl.min_t, aux_traj=True, noise_scale=wl.noise_scale,
                                  inpainting=True, input_aatype=True)
            return feats, noise, out, time.time() - t0

        feats, noise, out8, dt8 = run(8)
        _, _, out1, dt1 = run(1)
        torch.set_num_threads(8)

        def bb_rmsd(a, b):  # per-residue RMSD over N, CA, C, O (atom37 slots 0,1,2,4), unaligned
            d = a[..., [0, 1, 2, 4], :] - b[..., [0, 1, 2, 4], :]
            return np.sqrt((d ** 2).sum(-1).mean(-1))

        floor = bb_rmsd(out8["prot_traj"][0], out1["prot_traj"][0])
        print(f"{tag}: ref 8thr {dt8:.1f}s, 1thr {dt1:.1f}s, self-divergence floor max {floor.max():.3e} mean {floor.mean():.3e}")
        save = {f"in_{k}": v.numpy() for k, v in feats.items()}
        save.update(noise=noise, prot_traj=out8["prot_traj"][:, :, :, :5].astype(np.float32), rigid_traj=out8["rigid_traj"],
                    rigid_0_traj=out8["rigid_0_traj"][:, :, :, :5].astype(np.float32), trans_traj=out8["trans_traj"], psi_pred=out8["psi_pred"],
                    floor_8v1_final=floor, ref_seconds=np.array([dt8, dt1]))
        np.savez_compressed(os.path.join(OUT, tag + ".npz"), **save)


if __name__ == "__main__":
    main()
