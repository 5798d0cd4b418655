"""GPU parity on the configurations the metric is quoted on (BASELINE.json configs #2-#5), against fixtures written by the UNMODIFIED
reference (oracle/make_golden_large.py -> tests/golden/large_*.npz).  SURVEY.md §8c parity protocol: forward level (step 2),
teacher-forced step level (step 3), free-running trajectories (step 4).  Tolerances are written next to each assertion."""
import os

import numpy as np
import pytest
import torch

from util import bb_rmsd, rot_angle_between

pytestmark = pytest.mark.gpu

from framedipt_b200.config import default_conf  # noqa: E402


def _load(golden_dir, name):
    return dict(np.load(os.path.join(golden_dir, name)))


def _feats(g, device=None):
    f = {k[3:]: torch.tensor(v) for k, v in g.items() if k.startswith("in_")}
    if device is not None:
        f = {k: v.to(device) for k, v in f.items()}
    return f


def _model(de_novo=False, final_std=0.002):
    from framedipt_b200 import SE3Diffuser
    from framedipt_b200.params import synthetic_state_dict
    from framedipt_b200.score_network import ScoreNetwork

    conf = default_conf(input_aatype=not de_novo)
    diffuser = SE3Diffuser(conf.diffuser)
    m = ScoreNetwork(conf.model, diffuser, inpainting=not de_novo)
    m.load_state_dict(synthetic_state_dict(0, with_aatype=not de_novo, final_std=final_std))
    return m.to("cuda").eval(), diffuser


def _regen_noise(g, wl, diffuser, seed):
    """The legacy numpy stream the reference consumed: seed, the B sample_ref calls of make_features, then the trajectory's normals."""
    from framedipt_b200 import synthetic

    np.random.seed(123)
    feats = synthetic.make_features(wl, diffuser, seed=seed)
    noise = synthetic.draw_noise(wl.num_t, wl.batch, wl.n_res)
    chk = np.array([noise.sum(), np.abs(noise).sum(), noise.reshape(-1)[0], noise.reshape(-1)[-1]])
    assert np.allclose(chk, g["noise_checksum"], rtol=0, atol=1e-9), "regenerated noise differs from the stream the reference consumed"
    # x_T must be the reference's, too (same RNG consumption by sample_ref)
    assert np.abs(feats["rigids_t"].numpy()[..., 4:] - g["in_rigids_t"][..., 4:]).max() < 1e-5
    return noise


def _large_wl(tag):
    from framedipt_b200.synthetic import Workload

    return {
        "traj350": (Workload("traj350", 2, (170, 180), ((90, 102), (260, 272)), 100), 31),
        "traj500": (Workload("traj500", 1, (64,), ((20, 32),), 500), 32),
        "traj256dn": (Workload("traj256dn", 2, (256,), (), 50, de_novo=True), 33),
        "stress002": (Workload("stress002", 1, (64,), ((20, 32),), 50), 0),
    }[tag]


@pytest.mark.parametrize("tag,de_novo", [("traj350", False), ("traj256dn", True), ("traj500", False)])
def test_free_running_trajectory_vs_reference(golden_dir, tag, de_novo):
    """Free-running sampling on the bench geometries with the reference's noise stream: per-residue backbone RMSD of the final sample
    <= 1e-3 A (north star).  traj350 = cfg2 geometry (N=350, two chains, 100 steps); traj256dn = cfg3 geometry (de novo, N=256,
    50 steps); traj500 = a full 500-step schedule at N=64 with the reference's own 8-vs-1-thread floor printed beside it."""
    from framedipt_b200.inference import inference_fn

    g = _load(golden_dir, f"large_{tag}.npz")
    wl, seed = _large_wl(tag)
    m, diffuser = _model(de_novo)
    noise = _regen_noise(g, wl, diffuser, seed)
    out = inference_fn(m, diffuser, _feats(g, "cuda"), num_t=wl.num_t, min_t=wl.min_t, aux_traj=True, noise_scale=wl.noise_scale,
                       inpainting=not de_novo, input_aatype=not de_novo, noise=noise)
    slots = g["slots"]
    r_by_slot = [bb_rmsd(out["prot_traj"][s][:, :, :5], g["prot_traj"][k]).max() for k, s in enumerate(slots)]
    r = bb_rmsd(out["prot_traj"][0][:, :, :5], g["prot_traj"][0])
    floor = f"; reference 8-vs-1-thread floor {g['floor_8v1_final'].max():.3e}" if "floor_8v1_final" in g else ""
    print(f"{tag}: B={wl.batch} N={wl.n_res} {wl.num_t} steps: final per-residue RMSD vs reference max {r.max():.3e} mean {r.mean():.3e}{floor}; "
          f"by slot {dict(zip(slots.tolist(), [f'{x:.1e}' for x in r_by_slot]))}")
    assert r.max() < 1e-3
    assert r_by_slot[-1] < 2e-4  # first step (one forward + one reverse step)
    assert np.abs(out["trans_traj"][0] - g["trans_traj"][0]).max() < 1e-3
    assert np.abs(out["psi_pred"] - g["psi_pred"]).max() < 2e-3


@pytest.mark.parametrize("tag,de_novo", [("fwd129", False), ("fwd800", False), ("fwd1024", True)])
def test_forward_large_vs_reference(golden_dir, tag, de_novo):
    """One forward at N=129 (two j-tiles), N=800 (cfg4: four chains, ring-streaming IPA) and N=1024 (cfg5: the largest supported
    length) against the unmodified reference.  The pair side runs with fp16 operands (TF32 class), so the single-forward bound is the
    looser one of SURVEY §8c step 2 (5e-4 A / rad)."""
    g = _load(golden_dir, f"large_{tag}.npz")
    m, _ = _model(de_novo)
    out = m(_feats(g, "cuda"))
    r, r_ref = out["rigids"].cpu().numpy(), g["out_rigids"]
    dt = np.abs(r[..., 4:] - r_ref[..., 4:]).max()
    da = rot_angle_between(r[..., :4], r_ref[..., :4]).max()
    dts = np.abs(out["trans_score"].cpu().numpy() - g["out_trans_score"]).max() / np.abs(g["out_trans_score"]).max()
    drs = np.abs(out["rot_score"].cpu().numpy() - g["out_rot_score"]).max() / np.abs(g["out_rot_score"]).max()
    dpsi = np.abs(out["psi"].cpu().numpy() - g["out_psi"]).max()
    dbb = np.abs(out["atom37"].cpu().numpy()[:, :, :5] - g["out_atom37"]).max()
    print(f"{tag}: |dtrans| {dt:.2e} A, rot {da:.2e} rad, trans_score rel {dts:.2e}, rot_score rel {drs:.2e}, psi {dpsi:.2e}, atoms {dbb:.2e} A")
    assert dt < 5e-4 and da < 5e-4 and dts < 5e-4 and dpsi < 2e-3 and dbb < 1e-3
    assert drs < 2e-3  # float32 sin/cos of the 1000-term series (libm vs CUDA), see test_scores_grid


def test_seq_transformer_vs_reference_tap(golden_dir):
    """The sequence-transformer sub-block (ipa_pytorch.py:533-539) on the reference's own input: encoder output [B,N,320] against the
    forward hook on `seq_tfmr_0` of the unmodified reference at N=129 (two row tiles, ragged last tile)."""
    g = _load(golden_dir, "large_fwd129.npz")
    m, _ = _model(False)
    ctx = m.context(torch.device("cuda", 0))
    x_in = torch.tensor(g["tap_seq_tfmr_0_in"]).cuda()         # cat[node, skip_embed(init_node)]
    mask = torch.tensor(g["in_res_mask"], dtype=torch.float32).cuda()
    node0 = (torch.tensor(g["tap_node_embed_raw"]).cuda() * mask[..., None]).contiguous()
    tf, _ = ctx.seq_tfmr(0, x_in[..., :256].contiguous(), node0, mask.contiguous())
    ref = g["tap_seq_tfmr_0"]
    err = np.abs(tf.cpu().numpy() - ref).max()
    print(f"seq_tfmr_0: max abs err {err:.3e} (|ref| max {np.abs(ref).max():.2f})")
    assert err < 2e-5 * max(1.0, np.abs(ref).max())  # LayerNorm outputs of O(1): fp32-class


def test_teacher_forced_steps_vs_reference(golden_dir):
    """SURVEY §8c step 3 (experiments/utils.py:292-412): at every step of the reference's config-#1 trajectory, feed the REFERENCE's
    state (x_t, self-conditioning CA) and the same noise into forward -> reverse -> backbone and compare the atoms of x_{t-1} and of
    the x0 prediction: <= 5e-4 A (TF32-class pair side)."""
    from framedipt_b200.inference import build_schedule

    g = _load(golden_dir, "traj_cfg1.npz")
    m, diffuser = _model(False)
    dev = torch.device("cuda", 0)
    ctx = m.context(dev)
    feats = _feats(g, "cuda")
    T = g["prot_traj"].shape[0]
    _, sched, temb = build_schedule(diffuser, T, 0.01, 0.1)
    noise = torch.tensor(g["noise"]).cuda()
    worst, worst0 = 0.0, 0.0
    for s in range(1, T - 1):  # step s consumes the state after step s-1 = rigid_traj[T-s] and the x0 prediction of step s-1
        feats["rigids_t"] = torch.tensor(g["rigid_traj"][T - s]).cuda()
        feats["sc_ca_t"] = torch.tensor(g["rigid_0_traj"][T - s][:, :, 1]).cuda()  # CA of the previous x0 prediction
        pf = m.prepare(feats, dev)
        sc = sched[s:s + 1].copy()
        out = ctx.sample(pf, sc, temb[s:s + 1].contiguous(), noise[s:s + 1].contiguous(), self_condition=False)
        r = bb_rmsd(out["prot_traj"][0].cpu().numpy(), g["prot_traj"][T - 1 - s]).max()
        r0 = bb_rmsd(out["rigid_0_traj"][0].cpu().numpy(), g["rigid_0_traj"][T - 1 - s]).max()
        worst, worst0 = max(worst, r), max(worst0, r0)
    print(f"teacher-forced, {T - 2} steps: worst per-residue RMSD of x_(t-1) {worst:.3e} A, of the x0 prediction {worst0:.3e} A")
    assert worst < 5e-4 and worst0 < 5e-4


def test_stress_weights_vs_reference_floor(golden_dir):
    """Ill-conditioned stress case (zero-init layers <- N(0, 0.02)): the reference diverges from ITSELF by 0.37 A when
    only its CPU thread count changes (chaotic amplification of 1.5e-5 A reduction-order noise per forward).  The bar for such setups
    is "no worse than the reference's own floor": early steps agree tightly, and the final deviation from the reference is within 3x
    the reference's 8-vs-1-thread deviation.  Our per-forward deviation is larger than thread-order noise (TF32-class pair side,
    <= 5e-4 A), so the exponential growth starts from a higher level and crosses 1e-3 A a few steps earlier; both curves are printed."""
    from framedipt_b200.inference import inference_fn

    g = _load(golden_dir, "large_stress002.npz")
    wl, seed = _large_wl("stress002")
    m, diffuser = _model(False, final_std=0.02)
    noise = _regen_noise(g, wl, diffuser, seed)
    out = inference_fn(m, diffuser, _feats(g, "cuda"), num_t=wl.num_t, min_t=wl.min_t, aux_traj=True, noise_scale=wl.noise_scale,
                       inpainting=True, input_aatype=True, noise=noise)
    T = wl.num_t
    ours = np.array([bb_rmsd(out["prot_traj"][s][:, :, :5], g["prot_traj"][s]).max() for s in range(T)])  # slot order (0 = final)
    floor = g["floor_8v1_by_slot"]
    sl = [T - 1, T - 5, T - 10, T - 20, T - 30, 10, 5, 0]
    print("stress 0.02 (slot: ours / reference 8-vs-1-thread floor, A): " + ", ".join(f"{s}: {ours[s]:.1e}/{floor[s]:.1e}" for s in sl))
    assert ours[T - 1] < 5e-4 and ours[T - 10] < 2e-3
    assert ours[0] <= max(3.0 * floor[0], 1e-3), (ours[0], floor[0])
