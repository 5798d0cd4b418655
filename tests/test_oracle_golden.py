"""Pins the oracle restatement (oracle/framedipt_oracle.py) against fixtures produced by the UNMODIFIED
reference (oracle/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest
import torch

from util import bb_rmsd, rot_angle_between

from oracle import framedipt_oracle as orc


def _load(golden_dir, name):
    return dict(np.load(os.path.join(golden_dir, name)))


def _feats(g):
    return {k[3:]: torch.tensor(v) for k, v in g.items() if k.startswith("in_")}


@pytest.mark.parametrize("name", ["forward_small.npz", "forward_small_padded.npz"])
def test_forward_matches_reference(golden_dir, state_dict, name):
    g = _load(golden_dir, name)
    feats = _feats(g)
    taps = {}
    with torch.no_grad():
        out = orc.score_network_forward(state_dict, feats, taps=taps)
    valid = feats["res_mask"].numpy().astype(bool)
    assert np.abs(taps["node_embed0"].numpy() - g["tap_node_embed_raw"] * valid[..., None])[valid].max() < 2e-4
    em = valid[:, :, None] & valid[:, None, :]
    assert np.abs(taps["edge_embed0"].numpy() - g["tap_edge_embed_raw"])[em].max() < 5e-4
    assert np.abs(taps["ipa_0"].numpy() - g["tap_ipa_0"] * valid[..., None])[valid].max() < 2e-4
    assert np.abs(taps["edge_0"].numpy() - g["tap_edge_transition_0"])[em].max() < 1e-3
    r, r_ref = out["rigids"].numpy(), g["out_rigids"]
    assert np.abs(r[..., 4:] - r_ref[..., 4:])[valid].max() < 1e-4
    assert rot_angle_between(r[..., :4], r_ref[..., :4])[valid].max() < 1e-4
    ts_scale = np.abs(g["out_trans_score"]).max()
    assert np.abs(out["trans_score"].numpy() - g["out_trans_score"]).max() < 1e-4 * ts_scale
    rs_scale = np.abs(g["out_rot_score"]).max()
    assert np.abs(out["rot_score"].numpy() - g["out_rot_score"]).max() < 2e-3 * rs_scale
    assert np.abs(out["psi"].numpy() - g["out_psi"])[valid].max() < 1e-4


def test_scores_grid(golden_dir):
    g = _load(golden_dir, "scores_grid.npz")
    t = torch.tensor(g["t"])
    q_t = torch.tensor(g["q_t"])
    q_id = torch.zeros_like(q_t)
    q_id[..., 0] = 1
    s = orc.rot_score(q_t, q_id, t).numpy()
    ref = g["rot_score_identity0"]
    assert s.dtype == np.float64
    assert np.abs(s - ref).max() <= 1e-5 * np.abs(ref).max()
    s = orc.rot_score(q_t, torch.tensor(g["q_0"]), t).numpy()
    ref = g["rot_score_random0"]
    assert np.abs(s - ref).max() <= 1e-5 * np.abs(ref).max()
    ts = orc.trans_score(torch.tensor(g["trans_t"]), torch.tensor(g["trans_0"]), t[:, None, None]).numpy()
    assert np.abs(ts - g["trans_score"]).max() <= 1e-6 * np.abs(g["trans_score"]).max()


def test_reverse_and_backbone(golden_dir):
    g = _load(golden_dir, "reverse_small.npz")
    for ci in range(3):
        t, dt, ns, center = g[f"c{ci}_params"]
        R, T = orc.reverse_step(g["rigids_t"], g["rot_score"], g["trans_score"], g["diffuse_mask"], t, dt,
                                g[f"c{ci}_z_rot"], g[f"c{ci}_z_trans"], center=bool(center), noise_scale=ns)
        assert np.abs(R - g[f"c{ci}_rotmats"]).max() < 1e-6
        assert np.abs(T - g[f"c{ci}_trans"]).max() < 1e-5
    q = torch.tensor(g["rigids_t"][..., :4])
    bb = orc.compute_backbone(orc.quat_to_rot(q), g["rigids_t"][..., 4:], g["psi"], g["aatype"]).numpy()
    assert np.abs(bb - g["atom37"]).max() < 1e-5
    assert np.all(bb[..., 5:, :] == 0)


def test_sample_ref(golden_dir):
    from framedipt_b200 import synthetic

    g = _load(golden_dir, "sample_ref.npz")
    wl = synthetic.WORKLOADS["cfg1_monomer64"]
    st = synthetic.static_features(wl, 0)
    np.random.seed(123)
    r = orc.sample_ref(wl.n_res, st["gt_rotmats"], st["gt_trans"], 1.0 - st["fixed_mask"])
    ref = g["inpaint_rigids_t"][0]
    assert np.abs(r[:, 4:] - ref[:, 4:]).max() < 1e-5
    assert rot_angle_between(r[:, :4], ref[:, :4]).max() < 1e-3


def test_trajectory_small(golden_dir, state_dict):
    """Free-running 10-step trajectory (B=2, N=24) vs the reference with identical noise."""
    g = _load(golden_dir, "traj_small.npz")
    feats = _feats(g)
    out = orc.inference_loop(state_dict, feats, num_t=10, min_t=0.01, noise=g["noise"], noise_scale=0.1)
    r = bb_rmsd(out["prot_traj"][0], np.pad(g["prot_traj"][0], ((0, 0), (0, 0), (0, 32), (0, 0))))
    assert r.max() < 1e-3, r.max()
    assert out["prot_traj"].shape == (10, 2, 24, 37, 3)


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_trajectory_option_variants(golden_dir, state_dict, tag):
    """The oracle's center / noise_scale / self_condition / diffuse_rot / diffuse_trans switches against trajectories of the unmodified
    reference run with the same switches (oracle/make_golden_variants.py)."""
    g = _load(golden_dir, "traj_variants.npz")
    center, ns, sc, drot, dtrans = g[f"{tag}_opts"]
    feats = _feats(g)
    out = orc.inference_loop(state_dict, feats, num_t=int(g["num_t"]), min_t=0.01, noise=g[f"{tag}_noise"], noise_scale=float(ns),
                             center=bool(center), self_condition=bool(sc), diffuse_rot=bool(drot), diffuse_trans=bool(dtrans))
    r = bb_rmsd(out["prot_traj"][0][:, :, :5], g[f"{tag}_prot_traj"][0])
    assert r.max() < 1e-3, r.max()
    r_mid = bb_rmsd(out["prot_traj"][3][:, :, :5], g[f"{tag}_prot_traj"][3])
    assert r_mid.max() < 1e-3, r_mid.max()
