"""GPU parity tests: the CUDA path (through the C ABI of libfdpt.so) against the reference-generated golden
fixtures and the CPU oracle.  Tolerances are fp32-class (SURVEY §8c parity protocol)."""
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


@pytest.fixture(scope="module")
def model(state_dict):
    from framedipt_b200 import SE3Diffuser
    from framedipt_b200.score_network import ScoreNetwork

    conf = default_conf()
    diffuser = SE3Diffuser(conf.diffuser)
    m = ScoreNetwork(conf.model, diffuser, inpainting=True)
    m.load_state_dict(state_dict)
    m = m.to("cuda").eval()
    return m, diffuser


@pytest.fixture(scope="module")
def ctx(model):
    return model[0].context(torch.device("cuda", 0))


def test_linear(ctx):
    g = torch.Generator().manual_seed(1)
    for (M, N, K, act) in [(100, 70, 86, 1), (257, 256, 2688, 0), (64, 6, 256, 0)]:
        x, w, b = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) / K ** 0.5, torch.randn(N, generator=g)
        y = ctx.linear(x.cuda(), w.cuda(), b.cuda(), act).cpu()
        ref = torch.nn.functional.linear(x.double(), w.double(), b.double())
        if act:
            ref = ref.relu()
        # tensor-core fp32 accumulation truncates (RZ): split-TF32 is ~2x plain fp32 error, still fp32-class (single-pass TF32 is ~5e-4)
        assert (y - ref.float()).abs().max() < 1e-5 * max(1.0, ref.abs().max())


def test_tc_linear(ctx):
    """tcgen05 building block: fp16 operands (10-bit mantissa, same as TF32), fp32 accumulation."""
    g = torch.Generator().manual_seed(2)
    for (M, N, K, act) in [(128, 128, 64, 0), (300, 384, 128, 1), (1000, 384, 384, 1), (257, 128, 512, 0)]:
        x, w, b = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) / K ** 0.5, torch.randn(N, generator=g)
        y = ctx.tc_linear(x.cuda(), w.cuda(), b.cuda(), act).cpu()
        ref = torch.nn.functional.linear(x.half().double(), w.half().double(), b.double())  # exact product of the rounded operands
        if act:
            ref = ref.relu()
        err = (y - ref.float()).abs().max().item()
        assert err < 2e-5 * max(1.0, ref.abs().max().item()), (M, N, K, err)
        full = torch.nn.functional.linear(x.double(), w.double(), b.double())
        if act:
            full = full.relu()
        assert (y - full.float()).abs().max().item() < 5e-3 * max(1.0, full.abs().max().item())


@pytest.mark.parametrize("name", ["forward_small.npz", "forward_small_padded.npz"])
def test_embed_ipa_edge_kernels(golden_dir, model, ctx, name):
    g = _load(golden_dir, name)
    m, _ = model
    feats = _feats(g, "cuda")
    pf = m.prepare(feats, torch.device("cuda", 0))
    node, edge = ctx.embed(pf, feats["t"])
    valid = g["in_res_mask"].astype(bool)
    em = valid[:, :, None] & valid[:, None, :]
    assert np.abs(node.cpu().numpy() - g["tap_node_embed_raw"])[valid].max() < 2e-4
    # fused tcgen05 edge embedder: fp16 operands (TF32-class mantissa) through three chained GEMMs, fp32 accumulation, LayerNorm in
    # fp32, z stored as fp16 (|z| <~ 11 after LayerNorm) -- same error class as the fused EdgeTransition below
    e_ref = g["tap_edge_embed_raw"]
    e_err = np.abs(edge.cpu().numpy() - e_ref)[em]
    print(f"edge embed: max err {e_err.max():.3e} mean {e_err.mean():.3e}")
    assert e_err.max() < 2e-2 and e_err.mean() < 2e-3, (e_err.max(), e_err.mean())
    assert np.all(edge.cpu().numpy()[~em] == 0)
    # IPA block 0 on the reference's own inputs
    mask = torch.tensor(g["in_res_mask"], dtype=torch.float32).cuda()
    s = torch.tensor(g["tap_node_embed_raw"]).cuda() * mask[..., None]
    z = torch.tensor(g["tap_edge_embed_raw"]).cuda() * (mask[:, :, None] * mask[:, None, :])[..., None]
    rig = torch.tensor(g["in_rigids_t"]).cuda().float()
    out = ctx.ipa(0, s.contiguous(), z.contiguous(), rig[..., :4].contiguous(), (rig[..., 4:] * 0.1).contiguous(), mask.contiguous())
    ref = g["tap_ipa_0"]
    assert np.abs(out.cpu().numpy() - ref)[valid].max() < 1e-4 * max(1.0, np.abs(ref).max())
    # edge transition block 0
    node_in = torch.tensor(g["tap_node_transition_0"]).cuda() * mask[..., None]
    zo = ctx.edge_transition(0, node_in.contiguous(), z.contiguous(), mask.contiguous())
    # fused tcgen05 EdgeTransition: fp16 operands (TF32-class mantissa), fp32 accumulation, fp16 storage of z
    z_ref = g["tap_edge_transition_0"]
    err = np.abs(zo.cpu().numpy() - z_ref)[em]
    assert err.max() < 2e-2 and err.mean() < 2e-3, (err.max(), err.mean())
    assert np.all(zo.cpu().numpy()[~em] == 0)


@pytest.mark.parametrize("name", ["forward_small.npz", "forward_small_padded.npz"])
def test_forward_vs_reference(golden_dir, model, name):
    g = _load(golden_dir, name)
    m, _ = model
    out = m(_feats(g, "cuda"))
    valid = g["in_res_mask"].astype(bool)
    r, r_ref = out["rigids"].cpu().numpy(), g["out_rigids"]
    assert np.abs(r[..., 4:] - r_ref[..., 4:])[valid].max() < 1e-4  # Angstrom
    assert rot_angle_between(r[..., :4], r_ref[..., :4])[valid].max() < 1e-4  # rad
    ts = np.abs(g["out_trans_score"]).max()
    assert np.abs(out["trans_score"].cpu().numpy() - g["out_trans_score"]).max() < 1e-4 * ts
    rs = np.abs(g["out_rot_score"]).max()
    assert out["rot_score"].dtype == torch.float64
    assert np.abs(out["rot_score"].cpu().numpy() - g["out_rot_score"]).max() < 2e-3 * rs
    assert np.abs(out["psi"].cpu().numpy() - g["out_psi"])[valid].max() < 1e-4
    assert np.abs(out["atom37"].cpu().numpy()[:, :, :5] - g["out_atom37"])[valid].max() < 2e-4
    assert out["atom37"].shape == (2, 24, 37, 3) and out["atom14"].shape == (2, 24, 14, 3)


def test_scores_grid(golden_dir, model, ctx):
    g = _load(golden_dir, "scores_grid.npz")
    _, diffuser = model
    sigma = torch.tensor(diffuser._so3_diffuser.grid_sigma(g["t"])).cuda()
    q_t = torch.tensor(g["q_t"]).cuda()
    q_id = torch.zeros_like(q_t)
    q_id[..., 0] = 1
    from oracle import framedipt_oracle as orc

    for q0, key in ((q_id, "rot_score_identity0"), (torch.tensor(g["q_0"]).cuda(), "rot_score_random0")):
        s = ctx.rot_score(q_t, q0.contiguous(), sigma).cpu().numpy()
        ref = g[key]
        assert np.isfinite(s).all()
        # The reference evaluates sin/cos((l+1/2) omega) of its 1000-term series in float32; where the series cancels
        # massively (omega >> sigma: negligible density; omega -> 0 or pi) its own value is libm-dependent rounding noise
        # (see oracle.rot_score_conditioning).  Parity is asserted on the well-conditioned region the sampler lives in:
        # 0.01 <= omega <= min(4 sigma, 3.0).
        v = orc.quat_to_rotvec(orc.quat_multiply((q0.cpu() * torch.tensor([1.0, -1, -1, -1])), q_t.cpu()))
        om = torch.linalg.norm(v, dim=-1).numpy()
        sg = sigma.cpu().numpy()[:, None]
        good = (om >= 0.01) & (om <= np.minimum(4 * sg, 3.0))
        assert good.mean() > 0.4
        err = np.abs(s - ref).max(-1)
        scale = np.abs(ref).max(-1)
        assert (err[good] <= 1e-4 * scale[good]).all(), (err[good] / scale[good]).max()
    ts = ctx.trans_score(torch.tensor(g["trans_t"]).cuda(), torch.tensor(g["trans_0"]).cuda(), torch.tensor(g["t"]).cuda()).cpu().numpy()
    assert np.abs(ts - g["trans_score"]).max() <= 2e-6 * np.abs(g["trans_score"]).max()


def test_reverse_and_backbone(golden_dir, model, ctx):
    g = _load(golden_dir, "reverse_small.npz")
    _, diffuser = model
    rig = torch.tensor(g["rigids_t"]).cuda()
    for ci in range(3):
        t, dt, ns, center = g[f"c{ci}_params"]
        row = diffuser.step_scalars(float(t), float(dt), float(ns))
        out = ctx.reverse(rig, torch.tensor(g["rot_score"]).cuda(), torch.tensor(g["trans_score"]).cuda(),
                          torch.tensor(g["diffuse_mask"], dtype=torch.float32).cuda(), torch.tensor(g[f"c{ci}_z_rot"]).cuda(),
                          torch.tensor(g[f"c{ci}_z_trans"]).cuda(), row, center=bool(center)).cpu().numpy()
        assert np.abs(out[..., 4:] - g[f"c{ci}_trans"]).max() < 1e-5
        assert rot_angle_between(out[..., :4], g[f"c{ci}_tensor7"][..., :4]).max() < 2e-3  # fp32 quaternion of an fp32 matrix
        from oracle.framedipt_oracle import quat_to_rot
        R = quat_to_rot(torch.tensor(out[..., :4])).numpy()
        assert np.abs(R - g[f"c{ci}_rotmats"]).max() < 2e-6
    bb = ctx.backbone(rig, torch.tensor(g["psi"]).cuda(), torch.tensor(g["aatype"], dtype=torch.int32).cuda()).cpu().numpy()
    assert np.abs(bb - g["atom37"][:, :, :5]).max() < 1e-5


@pytest.mark.parametrize("name,num_t", [("traj_small.npz", 10), ("traj_cfg1.npz", 50)])
def test_trajectory_vs_reference(golden_dir, model, name, num_t):
    """Free-running sampling with the reference's noise: per-residue backbone RMSD <= 1e-3 Å (north star)."""
    from framedipt_b200.inference import inference_fn

    g = _load(golden_dir, name)
    m, diffuser = model
    out = inference_fn(m, diffuser, _feats(g, "cuda"), num_t=num_t, min_t=0.01, aux_traj=True, noise_scale=0.1, inpainting=True,
                       input_aatype=True, noise=g["noise"])
    T, B, N = g["prot_traj"].shape[:3]
    assert out["prot_traj"].shape == (T, B, N, 37, 3) and out["rigid_traj"].shape == (T + 1, B, N, 7)
    assert out["trans_traj"].shape == (T, B, N, 3) and out["psi_pred"].shape == (1, B, N, 2) and out["rigid_0_traj"].shape == (T, B, N, 37, 3)
    r = bb_rmsd(out["prot_traj"][0][:, :, :5], g["prot_traj"][0])
    print(f"{name}: per-residue RMSD max {r.max():.3e} mean {r.mean():.3e}; reference 8-vs-1-thread floor {g['floor_8v1_final'].max():.3e}")
    assert r.max() < 1e-3
    assert np.all(out["prot_traj"][..., 5:, :] == 0)
    # first reverse step (one forward) must agree tightly
    r_first = bb_rmsd(out["prot_traj"][-1][:, :, :5], g["prot_traj"][-1])
    assert r_first.max() < 2e-4
    assert np.abs(out["trans_traj"][0] - g["trans_traj"][0]).max() < 1e-3
    assert np.abs(out["rigid_traj"][-1] - g["rigid_traj"][-1]).max() == 0  # x_T passthrough


def test_no_cpu_fallback(model):
    m, _ = model
    with pytest.raises(Exception):
        m({"rigids_t": torch.zeros(1, 4, 7)})


def test_matmul_split_tf32(ctx):
    """Node-side GEMM kernel (tcgen05, 3-term split TF32): fp32-class accuracy, both B layouts, ragged sizes, batch strides."""
    g = torch.Generator().manual_seed(3)
    for (Bt, M, N, K, kmajor) in [(1, 128, 128, 32, True), (3, 350, 350, 280, True), (2, 257, 70, 86, True), (1, 1000, 960, 320, True),
                                  (2, 350, 292, 350, False), (1, 300, 80, 301, False), (1, 64, 256, 2688, True)]:
        a = torch.randn(Bt, M, K, generator=g).cuda()
        b = torch.randn(Bt, N, K, generator=g).cuda() if kmajor else torch.randn(Bt, K, N, generator=g).cuda()
        c = ctx.matmul(a, b, kmajor)
        ref = a.double() @ (b.double().transpose(1, 2) if kmajor else b.double())
        err = (c.double() - ref).abs().max().item()
        # fp32-class: the SIMT fp32 kernel measures ~7e-7 of max|C| on these shapes, split-TF32 ~1.2e-6 (the tensor core's fp32
        # accumulation truncates); single-pass TF32 would be ~5e-4
        # ... and its bias grows linearly with the number of k-steps: measured 7.7e-6 at K=2688 (linear_out)
        assert err < (2e-6 + 3e-9 * K) * ref.abs().max().item(), (Bt, M, N, K, kmajor, err)


def _oracle_forward_case(wl_args, de_novo, seed, t_val):
    """CUDA forward vs the CPU oracle (itself pinned to the reference fixtures) on a synthetic workload: covers shapes the committed
    fixtures cannot (N > 128: several j-tiles per row, the ring-streaming mode of the IPA kernel, ragged last tiles; the de-novo
    parameterisation without aatype)."""
    from framedipt_b200 import SE3Diffuser, synthetic
    from framedipt_b200.params import synthetic_state_dict
    from framedipt_b200.score_network import ScoreNetwork
    from oracle import framedipt_oracle as orc

    conf = default_conf(input_aatype=not de_novo)
    diffuser = SE3Diffuser(conf.diffuser)
    sd = synthetic_state_dict(0, with_aatype=not de_novo)
    m = ScoreNetwork(conf.model, diffuser, inpainting=not de_novo)
    m.load_state_dict(sd)
    m = m.to("cuda").eval()
    wl = synthetic.Workload(*wl_args, de_novo=de_novo)
    np.random.seed(seed)
    feats = synthetic.make_features(wl, diffuser, seed=seed)
    feats["t"] = t_val * torch.ones(wl.batch)
    feats["sc_ca_t"] = feats["rigids_t"][..., 4:].float() + 0.3 * torch.randn(wl.batch, wl.n_res, 3, generator=torch.Generator().manual_seed(seed))
    out = m({k: v.to("cuda") for k, v in feats.items()})
    torch.set_num_threads(os.cpu_count() or 1)
    with torch.no_grad():
        ref = orc.score_network_forward(sd, feats, inpainting=not de_novo, input_aatype=not de_novo)
    r, r_ref = out["rigids"].cpu().numpy(), ref["rigids"].float().numpy()
    dt = np.abs(r[..., 4:] - r_ref[..., 4:]).max()
    da = rot_angle_between(r[..., :4], r_ref[..., :4]).max()
    ts_ref = ref["trans_score"].float().numpy()
    dts = np.abs(out["trans_score"].cpu().numpy() - ts_ref).max() / np.abs(ts_ref).max()
    dpsi = np.abs(out["psi"].cpu().numpy() - ref["psi"].float().numpy()).max()
    print(f"N={wl.n_res} B={wl.batch} de_novo={de_novo}: |dtrans| {dt:.2e} A, rot {da:.2e} rad, trans_score rel {dts:.2e}, psi {dpsi:.2e}")
    # the pair side runs with fp16 operands (TF32-class), so the single-forward bound is the looser one of SURVEY §8c step 2
    # psi is a normalised 2-vector (ipa_pytorch.py:361-363): ill-conditioned where the raw torsion output is small, hence 2e-3
    assert dt < 5e-4 and da < 5e-4 and dts < 5e-4 and dpsi < 2e-3


def test_forward_vs_oracle_multi_tile():
    _oracle_forward_case(("mt300", 2, (140, 160), ((60, 72), (200, 212)), 10), False, 11, 0.6)


def test_forward_vs_oracle_ring_streaming():
    _oracle_forward_case(("mt700", 1, (300, 400), ((100, 112), (500, 512)), 10), False, 12, 0.3)


def test_forward_vs_oracle_de_novo():
    _oracle_forward_case(("dn150", 2, (150,), (), 10), True, 13, 0.8)


def test_long_trajectory_vs_oracle(model, state_dict):
    """200 free-running steps (N=64, config-#1 geometry) against the CPU oracle with the same noise stream: the divergence must stay
    inside the 1e-3 A budget over a long horizon too (the configs the metric is quoted on run 200-500 steps)."""
    from framedipt_b200 import synthetic
    from framedipt_b200.inference import inference_fn
    from oracle import framedipt_oracle as orc

    m, diffuser = model
    wl = synthetic.Workload("long64", 1, (64,), ((20, 32),), 200)
    np.random.seed(321)
    feats = synthetic.make_features(wl, diffuser, seed=5)
    noise = synthetic.draw_noise(wl.num_t, wl.batch, wl.n_res)
    out = inference_fn(m, diffuser, {k: v.to("cuda") for k, v in feats.items()}, num_t=wl.num_t, min_t=0.01, aux_traj=True, noise_scale=0.1,
                       inpainting=True, input_aatype=True, noise=noise)
    torch.set_num_threads(os.cpu_count() or 1)
    ref = orc.inference_loop(state_dict, feats, num_t=wl.num_t, min_t=0.01, noise=noise, noise_scale=0.1)
    r = bb_rmsd(out["prot_traj"][0][:, :, :5], ref["prot_traj"][0][:, :, :5])
    print(f"200-step trajectory vs oracle: per-residue RMSD max {r.max():.3e} mean {r.mean():.3e}")
    assert r.max() < 1e-3


def test_linear_shapes_cover_the_packed_kernel(ctx):
    """fdpt_linear runs the A-stationary lin_tc kernel (weights split on the fly for this unit entry): ragged M, tiny and wide N,
    K not a multiple of 64 / 8, every k-block count up to 5."""
    g = torch.Generator().manual_seed(5)
    for (M, N, K, act) in [(1, 6, 256, 0), (129, 2, 256, 0), (300, 128, 54, 0), (2800, 960, 320, 1), (77, 6816, 256, 0), (500, 64, 65, 1),
                           (128, 384, 128, 0), (1000, 320, 192, 1)]:
        x, w, b = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) / K ** 0.5, torch.randn(N, generator=g)
        y = ctx.linear(x.cuda(), w.cuda(), b.cuda(), act).cpu()
        ref = torch.nn.functional.linear(x.double(), w.double(), b.double())
        if act:
            ref = ref.relu()
        err = (y - ref.float()).abs().max().item()
        assert err < 1e-5 * max(1.0, ref.abs().max().item()), (M, N, K, err)


def test_tiny_and_oversized_inputs(model):
    """N = 5 (far below one tile) runs and matches the oracle; N > 1024 is rejected loudly (no fallback)."""
    from framedipt_b200 import runtime

    _oracle_forward_case(("tiny5", 1, (5,), ((1, 3),), 4), False, 21, 0.5)
    m, diffuser = model
    from framedipt_b200 import synthetic

    wl = synthetic.Workload("big", 1, (1100,), ((10, 20),), 4)
    np.random.seed(1)
    feats = synthetic.make_features(wl, diffuser, seed=1)
    feats["t"] = torch.ones(1)
    with pytest.raises(runtime.FdptError):
        m({k: v.to("cuda") for k, v in feats.items()})


@pytest.mark.parametrize("tag", ["s6", "s12"])
def test_logp_confidence_score_vs_reference(golden_dir, model, tag):
    """EigenFold confidence score (experiments/utils.py:752-869) with the synthetic network: forward-noise the sample with the
    reference's RNG stream, two GPU forwards per step, transition log-densities -- against the unmodified reference on CPU.
    The noised path is identical (same seed); the scores carry the network's TF32-class pair-side error, hence a bound of 1e-3 on the log-probabilities."""
    from framedipt_b200 import Rigid
    from framedipt_b200.inference import logp_confidence_score

    g = _load(golden_dir, "logp_small.npz")
    m, diffuser = model
    feats = _feats(g, "cuda")
    mask = ((1 - g["in_fixed_mask"]) * g["in_res_mask"])[0].astype(np.float64)
    rig0 = Rigid.from_tensor_7(torch.tensor(g["in_rigids_t"][0]))
    num_t, min_t = int(g[f"{tag}_args"][0]), float(g[f"{tag}_args"][1])
    np.random.seed(77)
    lp, lps = logp_confidence_score(m, diffuser, rig0, feats, mask, num_t, min_t, "cuda", True)
    ref, refs = float(g[f"{tag}_log_prob"]), g[f"{tag}_log_probs"]
    print(f"{tag}: log_prob {lp:.6f} vs reference {ref:.6f}; per-step max |diff| {np.abs(np.array(lps) - refs).max():.3e}")
    assert len(lps) == num_t
    assert np.abs(np.array(lps) - refs).max() < 1e-3  # absolute, on running sums of magnitude 5 .. 150 (measured 3e-5)


@pytest.mark.parametrize("center,noise_scale,self_condition", [(False, 1.0, True), (True, 0.5, False)])
def test_sampler_option_variants_vs_oracle(model, state_dict, center, noise_scale, self_condition):
    """The switches of `inference_fn` that change the arithmetic of the loop (experiments/utils.py:511-626): `center` (R3 reverse step
    re-centres the diffused residues or not), `noise_scale`, `self_condition` (pre-pass at t=1 or zeros) -- 12 free-running steps of a
    padded two-chain batch against the CPU oracle on the same noise stream."""
    from framedipt_b200 import synthetic
    from framedipt_b200.inference import inference_fn
    from oracle import framedipt_oracle as orc

    m, diffuser = model
    wl = synthetic.Workload("opt40", 2, (22, 18), ((4, 12), (26, 33)), 12)
    np.random.seed(99)
    feats = synthetic.make_features(wl, diffuser, seed=8)
    feats["res_mask"][1, -3:] = 0.0
    noise = synthetic.draw_noise(wl.num_t, wl.batch, wl.n_res)
    out = inference_fn(m, diffuser, {k: v.to("cuda") for k, v in feats.items()}, num_t=wl.num_t, min_t=0.01, center=center, aux_traj=True,
                       self_condition=self_condition, noise_scale=noise_scale, inpainting=True, input_aatype=True, noise=noise)
    torch.set_num_threads(os.cpu_count() or 1)
    ref = orc.inference_loop(state_dict, feats, num_t=wl.num_t, min_t=0.01, noise=noise, noise_scale=noise_scale, center=center,
                             self_condition=self_condition)
    valid = feats["res_mask"].numpy().astype(bool)
    r = bb_rmsd(out["prot_traj"][0][:, :, :5], ref["prot_traj"][0][:, :, :5])[valid]
    r_first = bb_rmsd(out["prot_traj"][-1][:, :, :5], ref["prot_traj"][-1][:, :, :5])[valid]
    print(f"center={center} noise_scale={noise_scale} self_condition={self_condition}: RMSD max {r.max():.3e} (first step {r_first.max():.3e})")
    assert r_first.max() < 2e-4 and r.max() < 1e-3


@pytest.mark.parametrize("diffuse_rot,diffuse_trans", [(False, True), (True, False)])
def test_partial_diffusion_vs_oracle(state_dict, diffuse_rot, diffuse_trans):
    """`diffuse_rot=False` / `diffuse_trans=False` (se3_diffuser.py:373-385): the undiffused component of every frame is passed
    through the reverse step unchanged; 8 free-running steps against the CPU oracle."""
    from framedipt_b200 import SE3Diffuser, synthetic
    from framedipt_b200.inference import inference_fn
    from framedipt_b200.score_network import ScoreNetwork
    from oracle import framedipt_oracle as orc

    conf = default_conf()
    conf.diffuser.diffuse_rot, conf.diffuser.diffuse_trans = diffuse_rot, diffuse_trans
    diffuser = SE3Diffuser(conf.diffuser)
    m = ScoreNetwork(conf.model, diffuser, inpainting=True)
    m.load_state_dict(state_dict)
    m = m.to("cuda").eval()
    wl = synthetic.Workload("part32", 1, (32,), ((6, 20),), 8)
    np.random.seed(17)
    feats = synthetic.make_features(wl, SE3Diffuser(default_conf().diffuser), seed=2)
    noise = synthetic.draw_noise(wl.num_t, wl.batch, wl.n_res)
    out = inference_fn(m, diffuser, {k: v.to("cuda") for k, v in feats.items()}, num_t=wl.num_t, min_t=0.01, aux_traj=True, noise_scale=0.3,
                       inpainting=True, input_aatype=True, noise=noise)
    torch.set_num_threads(os.cpu_count() or 1)
    ref = orc.inference_loop(state_dict, feats, num_t=wl.num_t, min_t=0.01, noise=noise, noise_scale=0.3, diffuse_rot=diffuse_rot,
                             diffuse_trans=diffuse_trans)
    r = bb_rmsd(out["prot_traj"][0][:, :, :5], ref["prot_traj"][0][:, :, :5])
    print(f"diffuse_rot={diffuse_rot} diffuse_trans={diffuse_trans}: RMSD max {r.max():.3e}")
    assert r.max() < 1e-3
    # the undiffused component really is untouched until the last step (which takes the predicted frames)
    rt = out["rigid_traj"]  # [T+1, B, N, 7], index 0 = final
    if not diffuse_trans:
        assert np.abs(rt[1, ..., 4:] - rt[-1, ..., 4:]).max() < 1e-6
    if not diffuse_rot:
        assert rot_angle_between(rt[1, ..., :4], rt[-1, ..., :4]).max() < 1e-5


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_trajectory_option_variants_vs_reference(golden_dir, state_dict, tag):
    """Same switches, CUDA path against the unmodified reference's own trajectories (tests/golden/traj_variants.npz):
    a = center off, noise_scale 1, no self-conditioning pre-pass; b = rotations not diffused; c = translations not diffused."""
    from framedipt_b200 import SE3Diffuser
    from framedipt_b200.inference import inference_fn
    from framedipt_b200.score_network import ScoreNetwork

    g = _load(golden_dir, "traj_variants.npz")
    center, ns, sc, drot, dtrans = g[f"{tag}_opts"]
    conf = default_conf()
    conf.diffuser.diffuse_rot, conf.diffuser.diffuse_trans = bool(drot), bool(dtrans)
    diffuser = SE3Diffuser(conf.diffuser)
    m = ScoreNetwork(conf.model, diffuser, inpainting=True)
    m.load_state_dict(state_dict)
    m = m.to("cuda").eval()
    out = inference_fn(m, diffuser, _feats(g, "cuda"), num_t=int(g["num_t"]), min_t=0.01, center=bool(center), aux_traj=True,
                       self_condition=bool(sc), noise_scale=float(ns), inpainting=True, input_aatype=True, noise=g[f"{tag}_noise"])
    r = bb_rmsd(out["prot_traj"][0][:, :, :5], g[f"{tag}_prot_traj"][0])
    print(f"variant {tag}: per-residue RMSD vs reference max {r.max():.3e}")
    assert r.max() < 1e-3


def test_ipa_operand_image_modes_agree(model):
    """The three ways the IPA attention GEMMs get their operands (FDPT_OPT_IPA_IMG): 1 = operand images written by the projection GEMM's
    epilogue (default), 2 = images written by a separate prep kernel, 0 = fp32 operands split on the fly — same arithmetic class
    (2-term fp16 split), so a forward at N=300 (three j-tiles per row, ragged last tile, two chains) must agree to fp32 noise."""
    from framedipt_b200 import synthetic

    m, diffuser = model
    ctx = m.context(torch.device("cuda", 0))
    wl = synthetic.Workload("img300", 2, (140, 160), ((60, 72), (200, 212)), 10)
    np.random.seed(4)
    feats = synthetic.make_features(wl, diffuser, seed=4)
    feats["t"] = 0.5 * torch.ones(wl.batch)
    feats["res_mask"][1, -5:] = 0.0
    feats = {k: v.to("cuda") for k, v in feats.items()}
    outs = {}
    try:
        for mode in (1, 2, 0):
            ctx.set_option(7, mode)
            outs[mode] = {k: v.cpu().numpy() for k, v in m(feats).items()}
        ctx.set_option(7, 1)
        ctx.set_option(8, 0)  # sequence-transformer attention on the fp32-operand GEMM path (FDPT_OPT_TF_IMG = 0)
        outs["tf0"] = {k: v.cpu().numpy() for k, v in m(feats).items()}
        ctx.set_option(8, 1)
        ctx.set_option(3, 65536)  # IPA core: the two-pass kernel (N > 384 path) instead of the single-pass one
        outs["two_pass"] = {k: v.cpu().numpy() for k, v in m(feats).items()}
    finally:
        ctx.set_option(7, 1)
        ctx.set_option(8, 1)
        ctx.set_option(3, 0)
    valid = feats["res_mask"].cpu().numpy().astype(bool)
    for mode in (2, 0, "tf0", "two_pass"):
        d = np.abs(outs[mode]["rigids"][..., 4:] - outs[1]["rigids"][..., 4:])[valid].max()
        da = rot_angle_between(outs[mode]["rigids"][..., :4], outs[1]["rigids"][..., :4])[valid].max()
        print(f"IPA image mode {mode} vs 1: |dtrans| {d:.2e} A, rot {da:.2e} rad")
        assert d < 2e-5 and da < 2e-5


def test_forward_with_lecun_scale_final_layers_vs_oracle():
    """all other parity cases use final-layer weights of N(0, 0.002).  Here every layer the reference zero-initialises gets the
    LeCun scale of a trained layer (std = 1/sqrt(fan_in) ~ 0.06, 30x larger), so residual updates, frames and pair activations are an
    order of magnitude larger: the fp16 operand images / fp16 z storage must neither saturate nor lose the forward-level agreement.
    (|z| stays <= ~11 by construction: every z is a LayerNorm output; the hidden EdgeTransition activations are the ones that grow.)"""
    from framedipt_b200 import SE3Diffuser, synthetic
    from framedipt_b200.params import synthetic_state_dict
    from framedipt_b200.score_network import ScoreNetwork
    from oracle import framedipt_oracle as orc

    conf = default_conf()
    diffuser = SE3Diffuser(conf.diffuser)
    sd = synthetic_state_dict(0, final_std=0.0625)
    m = ScoreNetwork(conf.model, diffuser, inpainting=True)
    m.load_state_dict(sd)
    m = m.to("cuda").eval()
    wl = synthetic.Workload("lecun150", 2, (70, 80), ((20, 32), (100, 112)), 10)
    np.random.seed(8)
    feats = synthetic.make_features(wl, diffuser, seed=8)
    feats["t"] = 0.4 * torch.ones(wl.batch)
    feats["sc_ca_t"] = feats["rigids_t"][..., 4:].float()
    out = m({k: v.to("cuda") for k, v in feats.items()})
    torch.set_num_threads(os.cpu_count() or 1)
    with torch.no_grad():
        ref = orc.score_network_forward(sd, feats, inpainting=True, input_aatype=True)
    r, r_ref = out["rigids"].cpu().numpy(), ref["rigids"].float().numpy()
    assert np.isfinite(r).all() and np.isfinite(out["trans_score"].cpu().numpy()).all()
    move = np.abs(r_ref[..., 4:] - feats["rigids_t"].numpy()[..., 4:]).max()
    dt = np.abs(r[..., 4:] - r_ref[..., 4:]).max()
    da = rot_angle_between(r[..., :4], r_ref[..., :4]).max()
    print(f"LeCun-scale final layers: frames move by up to {move:.1f} A in one forward; |dtrans| {dt:.2e} A, rot {da:.2e} rad")
    # TF32-class pair side: the forward-level deviation scales with the size of the update (5e-4 A at 1 A of motion)
    assert da < 5e-3 and dt < 1e-3 * max(move, 1.0)


def test_tcgen05_mma_with_a_operand_in_tmem(ctx):
    """tcgen05.mma kind::f16 with A read from tensor memory (lane = row, two fp16 K-elements per 32-bit column, written by tcgen05.st):
    the operand form the fused EdgeTransition kernel uses to feed r2 into its third GEMM without a shared-memory round trip."""
    g = torch.Generator().manual_seed(9)
    a, b = torch.randn(128, 64, generator=g).cuda(), torch.randn(128, 64, generator=g).cuda()
    d = ctx.tmem_a_selftest(a.contiguous(), b.contiguous()).cpu()
    ref = a.cpu().half().double() @ b.cpu().half().double().T
    err = (d.double() - ref).abs().max().item()
    print(f"TMEM-A MMA: max abs err {err:.3e} (|ref| max {ref.abs().max():.1f})")
    assert err < 2e-5 * ref.abs().max().item()
