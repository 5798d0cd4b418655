"""GPU parity of the configuration switches that change the arithmetic of the hot path (SURVEY.md §8b "Variant switches"), of the
padded mixed-length batch (§8 f4) and of the device-side x_T construction / device RNG (§8 f3), against fixtures written by the
UNMODIFIED reference (oracle/make_golden_config.py -> tests/golden/config_variants.npz)."""
import os

import numpy as np
import pytest
import torch

from util import bb_rmsd, rot_angle_between

pytestmark = pytest.mark.gpu

from framedipt_b200.config import default_conf  # noqa: E402


@pytest.fixture(scope="module")
def g(golden_dir):
    return dict(np.load(os.path.join(golden_dir, "config_variants.npz")))


def _feats(g, prefix, device="cuda"):
    return {k[len(prefix):]: torch.tensor(v).to(device) for k, v in g.items() if k.startswith(prefix)}


def _build(conf, sd, inpainting=True):
    from framedipt_b200 import SE3Diffuser
    from framedipt_b200.score_network import ScoreNetwork

    diffuser = SE3Diffuser(conf.diffuser)
    m = ScoreNetwork(conf.model, diffuser, inpainting=inpainting)
    m.load_state_dict(sd)
    return m.to("cuda").eval(), diffuser


def _lookup_mismatch(ours, ref, tol=1e-5):
    """fraction of entries of a table-look-up score that differ (a float32 omega on a bucket boundary may fall into the neighbour)"""
    scale = np.abs(ref).max(-1, keepdims=True) + 1e-30
    return float((np.abs(ours - ref) / scale > tol).any(-1).mean())


def test_cached_score_table_vs_reference(g, golden_dir, state_dict, tmp_path):
    """so3.use_cached_score=True (so3_diffuser.py:389-396): the IGSO(3) score norm is looked up in the [num_sigma, num_omega] table."""
    from framedipt_b200 import Rotation
    from framedipt_b200.inference import inference_fn

    conf = default_conf()
    conf.diffuser.so3.use_cached_score = True
    conf.diffuser.so3.cache_dir = str(tmp_path)
    m, diffuser = _build(conf, state_dict)
    grid = dict(np.load(os.path.join(golden_dir, "scores_grid.npz")))
    q_t, tt = torch.tensor(grid["q_t"]), torch.tensor(grid["t"])
    q_id = torch.zeros_like(q_t)
    q_id[..., 0] = 1
    from oracle import framedipt_oracle as orc

    for q0, key in ((q_id, "cached_rot_score_identity0"), (torch.tensor(grid["q_0"]), "cached_rot_score_random0")):
        s = diffuser.calc_rot_score(Rotation(quats=q_t), Rotation(quats=q0), tt).cpu().numpy()
        # same well-conditioned region as test_scores_grid (0.01 <= omega <= min(4 sigma, 3)): for omega -> 0 the float32 quaternion ->
        # rotation-vector conversion is rounding noise, and for omega >> sigma the table entries themselves are (values ~1e-9 where the
        # 1000-term series cancels; the host-built table and the reference's agree there to 6e-9 absolute, not relatively)
        v = orc.quat_to_rotvec(orc.quat_multiply((q0 * torch.tensor([1.0, -1, -1, -1])), q_t))
        om = torch.linalg.norm(v, dim=-1).numpy()
        sg = diffuser._so3_diffuser.grid_sigma(tt.numpy())[:, None]
        good = (om >= 0.01) & (om <= np.minimum(4 * sg, 3.0))
        bad = _lookup_mismatch(s[good], g[key][good])
        print(f"{key}: {bad:.3%} of the well-conditioned entries differ ({good.mean():.0%} of the grid)")
        assert np.isfinite(s).all() and good.mean() > 0.4 and bad < 0.02
    feats = _feats(g, "cached_in_")
    f1 = dict(feats)
    f1["t"] = torch.tensor([0.37, 0.81]).cuda()
    out = m(f1)
    assert np.abs(out["rigids"].cpu().numpy()[..., 4:] - g["cached_fwd_rigids"][..., 4:]).max() < 1e-4
    # inside a forward the rotation vector carries the network's error (1e-5 rad class) and most residues are fixed (omega ~ 1e-6: the
    # direction v / omega is rounding noise there): same criterion as the series score of the other forward tests -- error relative to the
    # largest score of the tensor -- with a few bucket flips of the table index allowed
    rs, rs_ref = out["rot_score"].cpu().numpy(), g["cached_fwd_rot_score"]
    frac_bad = float((np.abs(rs - rs_ref).max(-1) > 2e-3 * np.abs(rs_ref).max()).mean())
    print(f"use_cached_score forward: {frac_bad:.2%} of the residues off by more than 2e-3 of max|score|")
    assert frac_bad < 0.05
    traj = inference_fn(m, diffuser, feats, num_t=6, min_t=0.01, aux_traj=True, noise_scale=0.1, inpainting=True, input_aatype=True,
                        noise=g["cached_noise"])
    r = bb_rmsd(traj["prot_traj"][0][:, :, :5], g["cached_prot_traj"][0])
    print(f"use_cached_score: 6-step RMSD vs reference max {r.max():.3e}")
    assert r.max() < 1e-3


def test_no_self_conditioning_features_vs_reference(g):
    """model.embed.embed_self_conditioning=False: the edge embedder has no distogram input (score_network.py:95-96, 185) and
    inference_fn(embed_self_conditioning=False) neither runs the pre-pass nor updates sc_ca_t (experiments/utils.py:356-358, 571-578)."""
    from framedipt_b200.inference import inference_fn
    from framedipt_b200.params import ModelDims, synthetic_state_dict

    conf = default_conf()
    conf.model.embed.embed_self_conditioning = False
    sd = synthetic_state_dict(0, ModelDims(embed_self_conditioning=False))
    assert sd["embedding_layer.edge_embedder.0.weight"].shape == (128, 140)
    m, diffuser = _build(conf, sd)
    feats = _feats(g, "nosc_in_")
    f1 = dict(feats)
    f1["t"] = torch.tensor([0.37, 0.81]).cuda()
    f1["sc_ca_t"] = torch.tensor(g["nosc_fwd_sc_ca_t"]).cuda()
    out = m(f1)
    r = out["rigids"].cpu().numpy()
    assert np.abs(r[..., 4:] - g["nosc_fwd_rigids"][..., 4:]).max() < 1e-4
    assert rot_angle_between(r[..., :4], g["nosc_fwd_rigids"][..., :4]).max() < 1e-4
    assert np.abs(out["psi"].cpu().numpy() - g["nosc_fwd_psi"]).max() < 1e-4
    f1["sc_ca_t"] = torch.zeros_like(f1["sc_ca_t"])
    assert torch.equal(m(f1)["rigids"], out["rigids"])  # the self-conditioning input has no effect at all
    traj = inference_fn(m, diffuser, feats, num_t=6, min_t=0.01, aux_traj=True, noise_scale=0.1, embed_self_conditioning=False,
                        inpainting=True, input_aatype=True, noise=g["nosc_noise"])
    rr = bb_rmsd(traj["prot_traj"][0][:, :, :5], g["nosc_prot_traj"][0])
    print(f"embed_self_conditioning=False: 6-step RMSD vs reference max {rr.max():.3e}")
    assert rr.max() < 1e-3


@pytest.mark.parametrize("tag,inpainting,input_aatype", [("bbff", False, False), ("bbtf", True, False)])
def test_backbone_residue_types_follow_the_call_flags(g, state_dict, tag, inpainting, input_aatype):
    """inference_fn derives the residue types of the trajectory's backbone atoms from ITS OWN inpainting / input_aatype arguments
    (experiments/utils.py:549-555), not from the model's: (False, False) -> ALA frames everywhere (a GLY gets a CB);
    (True, False) -> diffused residues are 'unknown' (ALA), fixed ones keep their type (GLY has no CB)."""
    from framedipt_b200.inference import inference_fn

    m, diffuser = _build(default_conf(), state_dict)
    feats = _feats(g, "bb_in_")
    traj = inference_fn(m, diffuser, feats, num_t=4, min_t=0.01, aux_traj=True, noise_scale=0.1, inpainting=inpainting,
                        input_aatype=input_aatype, noise=g[f"{tag}_noise"])
    for key in ("prot_traj", "rigid_0_traj"):
        d = np.abs(traj[key][:, :, :, :5] - g[f"{tag}_{key}"]).max()
        print(f"{tag} {key}: max abs diff {d:.3e}")
        assert d < 1e-3
    cb_gly = np.abs(traj["prot_traj"][0, 0, 0, 3]).sum()  # residue 0 of sample 0 is GLY and fixed
    assert (cb_gly > 0) == (not inpainting)


def test_padded_mixed_length_batch_vs_reference(g, state_dict):
    """Two different structures (N=20 two chains, N=31) padded with pad_feats / pad_rigid (framedipt/data/utils.py:311-339) into one
    batch: valid residues must match the unmodified reference on the same padded batch, forward and 6-step trajectory."""
    from framedipt_b200.inference import inference_fn

    m, diffuser = _build(default_conf(), state_dict)
    feats = _feats(g, "mixed_in_")
    valid = g["mixed_in_res_mask"].astype(bool)
    f1 = dict(feats)
    f1["t"] = torch.tensor([0.6, 0.6]).cuda()
    f1["sc_ca_t"] = torch.tensor(g["mixed_fwd_sc_ca_t"]).cuda()
    out = m(f1)
    r = out["rigids"].cpu().numpy()
    dt = np.abs(r[..., 4:] - g["mixed_fwd_rigids"][..., 4:])[valid].max()
    da = rot_angle_between(r[..., :4], g["mixed_fwd_rigids"][..., :4])[valid].max()
    dts = np.abs(out["trans_score"].cpu().numpy() - g["mixed_fwd_trans_score"])[valid].max() / np.abs(g["mixed_fwd_trans_score"]).max()
    # and the padded sample equals the same structure run alone (the reference's nested-tensor key-padding semantics)
    d_alone = np.abs(r[0, :20, 4:] - g["mixA_alone_fwd_rigids"][0, :, 4:]).max()
    print(f"mixed batch forward: |dtrans| {dt:.2e} A, rot {da:.2e} rad, trans_score rel {dts:.2e}; padded vs alone {d_alone:.2e} A")
    assert dt < 1e-4 and da < 1e-4 and dts < 1e-4 and d_alone < 1e-4
    traj = inference_fn(m, diffuser, feats, num_t=6, min_t=0.01, aux_traj=True, noise_scale=0.1, inpainting=True, input_aatype=True,
                        noise=g["mixed_noise"])
    rr = bb_rmsd(traj["prot_traj"][0][:, :, :5], g["mixed_prot_traj"][0])[valid]
    print(f"mixed batch 6-step RMSD vs reference max {rr.max():.3e}")
    assert rr.max() < 1e-3


def test_device_sample_ref_matches_reference(golden_dir, model_ctx):
    """x_T built on the device (fdpt_sample_ref) from the reference's own draws = the fixture written by the unmodified
    SE3Diffuser.sample_ref (tests/golden/sample_ref.npz), inpainting and de novo."""
    from framedipt_b200 import SE3Diffuser, synthetic
    from framedipt_b200.sampler import sample_ref_batch

    ref = dict(np.load(os.path.join(golden_dir, "sample_ref.npz")))
    diffuser = SE3Diffuser(default_conf().diffuser)
    wl = synthetic.WORKLOADS["cfg1_monomer64"]
    st = synthetic.static_features(wl, 0)
    np.random.seed(123)
    x = sample_ref_batch(model_ctx, diffuser, 1, wl.n_res, gt_rotmats=st["gt_rotmats"], gt_trans=st["gt_trans"],
                         diffuse_mask=1.0 - st["fixed_mask"], rng="numpy").cpu().numpy()
    assert np.abs(x[..., 4:] - ref["inpaint_rigids_t"][..., 4:]).max() < 1e-5
    assert rot_angle_between(x[..., :4], ref["inpaint_rigids_t"][..., :4]).max() < 1e-5
    x = sample_ref_batch(model_ctx, diffuser, 2, 32, rng="numpy").cpu().numpy()  # the fixture's stream continues (no re-seed)
    assert np.abs(x[..., 4:] - ref["denovo_rigids_t"][..., 4:]).max() < 1e-5
    assert rot_angle_between(x[..., :4], ref["denovo_rigids_t"][..., :4]).max() < 1e-5
    # throughput mode: Philox draws -- right distribution (IGSO(3) angle cdf at t=1, N(0, 10 A) translations), reproducible, seed-dependent
    a = sample_ref_batch(model_ctx, diffuser, 64, 256, rng="philox", philox_seed=7).cpu().numpy()
    b = sample_ref_batch(model_ctx, diffuser, 64, 256, rng="philox", philox_seed=7).cpu().numpy()
    c = sample_ref_batch(model_ctx, diffuser, 64, 256, rng="philox", philox_seed=8).cpu().numpy()
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    tr = a[..., 4:].reshape(-1)
    assert abs(tr.mean()) < 0.1 and abs(tr.std() - 10.0) < 0.1
    ang = 2 * np.arccos(np.clip(np.abs(a[..., 0]), 0, 1)).reshape(-1)
    so3 = diffuser._so3_diffuser
    cdf = so3.cdf_row(int(so3.t_to_idx(1.0)))
    emp = np.searchsorted(np.sort(ang), so3.discrete_omega) / ang.size
    assert np.abs(emp - cdf / cdf[-1]).max() < 0.02  # Kolmogorov distance, 16k samples


def test_philox_throughput_mode(model_ctx, state_dict):
    """fdpt_sample with device-drawn normals (noise=NULL): reproducible for a seed, different across seeds, and statistically the same
    sampler -- the per-step translation increments of the diffused residues have the variance the schedule prescribes."""
    from framedipt_b200 import SE3Diffuser, synthetic
    from framedipt_b200.inference import inference_fn
    from framedipt_b200.score_network import ScoreNetwork

    conf = default_conf()
    diffuser = SE3Diffuser(conf.diffuser)
    m = ScoreNetwork(conf.model, diffuser, inpainting=True)
    m.load_state_dict(state_dict)
    m = m.to("cuda").eval()
    wl = synthetic.Workload("ph64", 4, (64,), ((8, 56),), 12)
    np.random.seed(5)
    feats = {k: v.cuda() for k, v in synthetic.make_features(wl, diffuser, seed=9).items()}
    kw = dict(num_t=wl.num_t, min_t=0.01, aux_traj=True, noise_scale=1.0, inpainting=True, input_aatype=True, rng="philox")
    a = inference_fn(m, diffuser, feats, philox_seed=11, **kw)
    b = inference_fn(m, diffuser, feats, philox_seed=11, **kw)
    c = inference_fn(m, diffuser, feats, philox_seed=12, **kw)
    assert np.array_equal(a["prot_traj"], b["prot_traj"]) and not np.array_equal(a["prot_traj"], c["prot_traj"])
    assert np.isfinite(a["prot_traj"]).all()
    # the numpy-stream run of the same problem: same network, different noise -> same order of magnitude of per-step motion
    np.random.seed(6)
    d = inference_fn(m, diffuser, feats, **{**kw, "rng": "numpy"})
    step_p = np.abs(np.diff(a["rigid_traj"][1:, :, 8:56, 4:], axis=0)).mean()
    step_n = np.abs(np.diff(d["rigid_traj"][1:, :, 8:56, 4:], axis=0)).mean()
    print(f"mean |dx| per step: philox {step_p:.3f} A, numpy stream {step_n:.3f} A")
    assert 0.8 < step_p / step_n < 1.25


@pytest.fixture(scope="module")
def model_ctx(state_dict):
    m, _ = _build(default_conf(), state_dict)
    return m.context(torch.device("cuda", 0))
