"""TEST INFRASTRUCTURE ONLY — CPU restatement of FrameDiPT's sampler hot path.

This module is the *oracle* (checker) for the CUDA product path in ``framedipt_b200/``.  It may be
imported only by ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / reference
arm.  It is an independent restatement (torch CPU for the network, numpy float64 for the reverse
SDE step exactly like the reference) written from the formulas of the reference; each function
cites the reference file:line it follows (paths relative to the reference repo root).

Parity pinning: the reference's own tests hold no golden vectors for this path, so the oracle is
pinned against outputs of the unmodified reference run in the build container
(``oracle/make_golden.py`` → ``tests/golden/*.npz``; checked by ``tests/test_oracle_golden.py``).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

import importlib.util
import os

_bc_path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "framedipt_b200", "backbone_constants.py")
_spec = importlib.util.spec_from_file_location("_fdpt_backbone_constants", _bc_path)
_bc = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_bc)

# default dims, config/base.yaml:55-79
H, C_HID, PQ, PV, C_S, C_Z, C_SKIP, N_BLOCKS = 8, 256, 8, 12, 256, 128, 64, 4
MIN_SIGMA, MAX_SIGMA, NUM_SIGMA = 0.1, 1.5, 1000  # base.yaml:46-50
MIN_B, MAX_B, COORD_SCALE = 0.1, 20.0, 0.1  # base.yaml:38-41


# --------------------------------------------------------------------------------------------
# small helpers
# --------------------------------------------------------------------------------------------
def linear(x, sd, name):
    return F.linear(x, sd[name + ".weight"].to(x.dtype), sd[name + ".bias"].to(x.dtype))


def layer_norm(x, sd, name):
    return F.layer_norm(x, (x.shape[-1],), sd[name + ".weight"].to(x.dtype), sd[name + ".bias"].to(x.dtype), 1e-5)


def quat_to_rot(q):
    """openfold/utils/rigid_utils.py:185-205 (no normalisation: assumes unit q)."""
    a, b, c, d = q.unbind(-1)
    return torch.stack([
        torch.stack([a * a + b * b - c * c - d * d, 2 * (b * c - a * d), 2 * (b * d + a * c)], -1),
        torch.stack([2 * (b * c + a * d), a * a - b * b + c * c - d * d, 2 * (c * d - a * b)], -1),
        torch.stack([2 * (b * d - a * c), 2 * (c * d + a * b), a * a - b * b - c * c + d * d], -1),
    ], -2)


def quat_multiply(q1, q2):
    """Hamilton product, rigid_utils.py:229-263."""
    a1, b1, c1, d1 = q1.unbind(-1)
    a2, b2, c2, d2 = q2.unbind(-1)
    return torch.stack([
        a1 * a2 - b1 * b2 - c1 * c2 - d1 * d2,
        a1 * b2 + b1 * a2 + c1 * d2 - d1 * c2,
        a1 * c2 - b1 * d2 + c1 * a2 + d1 * b2,
        a1 * d2 + b1 * c2 - c1 * b2 + d1 * a2,
    ], -1)


def rot_to_quat_np(R):
    """Unit quaternion (w,x,y,z) of a rotation matrix; stands in for rigid_utils.py:208-227
    (eigenvector of the 4x4 K matrix; sign arbitrary there). float64 numpy."""
    from scipy.spatial.transform import Rotation

    q = Rotation.from_matrix(R.reshape(-1, 3, 3)).as_quat()  # x,y,z,w
    q = np.concatenate([q[:, 3:], q[:, :3]], -1)
    return q.reshape(R.shape[:-2] + (4,))


# --------------------------------------------------------------------------------------------
# embeddings  (framedipt/model/score_network.py:17-64)
# --------------------------------------------------------------------------------------------
def index_embedding(indices, embed_size=32, max_len=2056):
    """score_network.py:17-38 — evaluated on the torch default path (int64 index -> fp32 argument)."""
    k = torch.arange(embed_size // 2)
    arg = indices[..., None] * math.pi / (max_len ** (2 * k[None] / embed_size))
    return torch.cat([torch.sin(arg), torch.cos(arg)], -1)


def timestep_embedding(t, dim=32, max_positions=10000):
    """score_network.py:41-64 (t: float32 [B])."""
    t = t * max_positions
    half = dim // 2
    emb = torch.exp(torch.arange(half, dtype=torch.float32) * -(math.log(max_positions) / (half - 1)))
    emb = t.float()[:, None] * emb[None, :]
    return torch.cat([torch.sin(emb), torch.cos(emb)], 1)


def calc_distogram(pos, min_bin=1e-5, max_bin=20.0, num_bins=22):
    """framedipt/data/utils.py:541-550."""
    d = torch.linalg.norm(pos[:, :, None, :] - pos[:, None, :, :], dim=-1)[..., None]
    lower = torch.linspace(min_bin, max_bin, num_bins)
    upper = torch.cat([lower[1:], lower.new_tensor([1e8])], -1)
    return ((d > lower) * (d < upper)).to(pos.dtype)


def preprocess_aatype(aatype, fixed_mask, inpainting, input_aatype):
    """framedipt/data/utils.py:565-610."""
    if aatype is None or (not inpainting and not input_aatype):
        return None
    aatype = aatype.to(torch.int64)
    if not input_aatype:
        aatype = torch.where(fixed_mask.bool(), aatype, torch.full_like(aatype, 20))
    return aatype


def embedder(sd, seq_idx, t, fixed_mask, sc_ca, aatype, dtype=torch.float32):
    """Embedder.forward, score_network.py:129-197. Returns node [B,N,256], edge [B,N,N,128]."""
    B, N = seq_idx.shape
    fm = fixed_mask[..., None]
    t_emb = timestep_embedding(t)[:, None, :].expand(B, N, -1)
    if aatype is not None:
        oh = F.one_hot(aatype, 21)
        eps_emb = timestep_embedding(torch.ones_like(t) * 1e-5)[:, None, :].expand(B, N, -1)
        comb = torch.where(fm.bool(), eps_emb, t_emb)
        prot = torch.cat([oh, comb, fm], -1)
    else:
        prot = torch.cat([t_emb, fm], -1)
    node_feats = torch.cat([prot, index_embedding(seq_idx)], -1).float()
    rel = (seq_idx[:, :, None] - seq_idx[:, None, :]).reshape(B, N * N)
    cross = torch.cat([prot[:, :, None, :].expand(B, N, N, -1), prot[:, None, :, :].expand(B, N, N, -1)], -1)
    pair = [cross.float().reshape(B, N * N, -1), index_embedding(rel)]
    pair.append(calc_distogram(sc_ca).reshape(B, N * N, -1))
    pair_feats = torch.cat(pair, -1).float()

    def mlp(x, pre):
        x = x.to(dtype)
        x = F.relu(linear(x, sd, pre + ".0"))
        x = F.relu(linear(x, sd, pre + ".2"))
        x = linear(x, sd, pre + ".4")
        return layer_norm(x, sd, pre + ".5")

    node = mlp(node_feats, "embedding_layer.node_embedder")
    edge = mlp(pair_feats, "embedding_layer.edge_embedder").reshape(B, N, N, -1)
    return node, edge


# --------------------------------------------------------------------------------------------
# trunk modules  (framedipt/model/ipa_pytorch.py)
# --------------------------------------------------------------------------------------------
def ipa(sd, pre, s, z, quats, trans, mask):
    """InvariantPointAttention.forward, ipa_pytorch.py:170-329. quats [B,N,4], trans (0.1 Å units) [B,N,3]."""
    B, N, _ = s.shape
    R = quat_to_rot(quats)  # [B,N,3,3]
    q = linear(s, sd, pre + ".linear_q").view(B, N, H, C_HID)
    kv = linear(s, sd, pre + ".linear_kv").view(B, N, H, 2 * C_HID)
    k, v = kv[..., :C_HID], kv[..., C_HID:]

    def points(name, npts):
        p = linear(s, sd, pre + name)  # [B,N,H*npts*3] laid out as x-block | y-block | z-block
        p = torch.stack(torch.split(p, p.shape[-1] // 3, dim=-1), -1)  # [B,N,H*npts,3]
        p = torch.einsum("bnij,bnpj->bnpi", R, p) + trans[:, :, None, :]
        return p.view(B, N, H, npts, 3)

    q_pts = points(".linear_q_points", PQ)
    kv_pts = points(".linear_kv_points", PQ + PV)
    k_pts, v_pts = kv_pts[..., :PQ, :], kv_pts[..., PQ:, :]

    b = linear(z, sd, pre + ".linear_b")  # [B,N,N,H]
    a = torch.einsum("bihc,bjhc->bhij", q, k) * math.sqrt(1.0 / (3 * C_HID))
    a = a + math.sqrt(1.0 / 3) * b.permute(0, 3, 1, 2)
    d2 = ((q_pts[:, :, None] - k_pts[:, None, :]) ** 2).sum(-1)  # [B,N,N,H,PQ]
    hw = F.softplus(sd[pre + ".head_weights"].to(s.dtype)) * math.sqrt(1.0 / (3 * (PQ * 9.0 / 2)))
    pt_att = (d2 * hw[None, None, None, :, None]).sum(-1) * (-0.5)  # [B,N,N,H]
    a = a + pt_att.permute(0, 3, 1, 2)
    sq_mask = 1e5 * (mask[:, :, None] * mask[:, None, :] - 1)
    a = torch.softmax(a + sq_mask[:, None], -1)  # [B,H,N,N]

    o = torch.einsum("bhij,bjhc->bihc", a, v).reshape(B, N, H * C_HID)
    o_pt = torch.einsum("bhij,bjhpx->bihpx", a, v_pts)  # global frame
    o_pt = torch.einsum("bnji,bnhpj->bnhpi", R, o_pt - trans[:, :, None, None, :])  # R^T (p - t)
    o_pt_norm = torch.sqrt((o_pt ** 2).sum(-1) + 1e-8).reshape(B, N, H * PV)
    o_pt = o_pt.reshape(B, N, H * PV, 3)
    pair_z = linear(z, sd, pre + ".down_z")  # [B,N,N,32]
    o_pair = torch.einsum("bhij,bijc->bihc", a, pair_z).reshape(B, N, H * (C_Z // 4))
    feats = torch.cat([o, o_pt[..., 0], o_pt[..., 1], o_pt[..., 2], o_pt_norm, o_pair], -1)
    return linear(feats, sd, pre + ".linear_out")


def seq_transformer(sd, pre, x, mask, nhead=4, nlayers=2):
    """torch.nn.TransformerEncoder(post-norm, ReLU, dropout 0) as configured at ipa_pytorch.py:433-443,
    with boolean key-padding semantics (the eval/no_grad fast path of torch)."""
    B, N, D = x.shape
    dh = D // nhead
    for l in range(nlayers):
        p = f"{pre}.layers.{l}"
        qkv = F.linear(x, sd[p + ".self_attn.in_proj_weight"].to(x.dtype), sd[p + ".self_attn.in_proj_bias"].to(x.dtype))
        q, k, v = qkv.split(D, -1)
        q = q.view(B, N, nhead, dh).transpose(1, 2)
        k = k.view(B, N, nhead, dh).transpose(1, 2)
        v = v.view(B, N, nhead, dh).transpose(1, 2)
        att = (q @ k.transpose(-1, -2)) / math.sqrt(dh)
        att = att.masked_fill(mask[:, None, None, :] == 0, float("-inf"))
        att = torch.softmax(att, -1)
        o = (att @ v).transpose(1, 2).reshape(B, N, D)
        o = linear(o, sd, p + ".self_attn.out_proj")
        x = layer_norm(x + o, sd, p + ".norm1")
        ff = linear(F.relu(linear(x, sd, p + ".linear1")), sd, p + ".linear2")
        x = layer_norm(x + ff, sd, p + ".norm2")
    return x


def node_transition(sd, pre, s):
    """StructureModuleTransition, ipa_pytorch.py:36-58."""
    s0 = s
    s = F.relu(linear(s, sd, pre + ".linear_1"))
    s = F.relu(linear(s, sd, pre + ".linear_2"))
    s = linear(s, sd, pre + ".linear_3")
    return layer_norm(s + s0, sd, pre + ".ln")


def edge_transition(sd, pre, node, z):
    """EdgeTransition.forward, ipa_pytorch.py:84-102."""
    B, N, _ = node.shape
    n = linear(node, sd, pre + ".initial_embed")
    x = torch.cat([z, n[:, :, None, :].expand(B, N, N, -1), n[:, None, :, :].expand(B, N, N, -1)], -1)
    y = F.relu(linear(x, sd, pre + ".trunk.0"))
    y = F.relu(linear(y, sd, pre + ".trunk.2"))
    y = linear(y + x, sd, pre + ".final_layer")
    return layer_norm(y, sd, pre + ".layer_norm")


def compose_q_update(quats, trans, upd, mask):
    """Rigid.compose_q_update_vec, rigid_utils.py:1039-1063 / 587-616 / 266-275."""
    qv = torch.cat([torch.zeros_like(upd[..., :1]), upd[..., :3]], -1)
    dq = quat_multiply(quats, qv)
    newq = quats + dq * mask
    newq = newq / torch.linalg.norm(newq, dim=-1, keepdim=True)
    R = quat_to_rot(quats)
    newt = trans + torch.einsum("bnij,bnj->bni", R, upd[..., 3:]) * mask
    return newq, newt


def torsion_angles(sd, pre, s):
    """TorsionAngles.forward, ipa_pytorch.py:347-363 (normalised output only)."""
    s0 = s
    s = F.relu(linear(s, sd, pre + ".linear_1"))
    s = linear(s, sd, pre + ".linear_2") + s0
    u = linear(s, sd, pre + ".linear_final")
    return u / torch.sqrt(torch.clamp((u ** 2).sum(-1, keepdim=True), min=1e-8))


# --------------------------------------------------------------------------------------------
# scores  (framedipt/diffusion/*)
# --------------------------------------------------------------------------------------------
def sigma_of_t(t):
    """so3_diffuser.py:299-306 (numpy float64)."""
    return np.log(t * np.exp(MAX_SIGMA) + (1 - t) * np.exp(MIN_SIGMA))


def discrete_sigma():
    return sigma_of_t(np.linspace(0.0, 1.0, NUM_SIGMA))


def sigma_grid_value(t):
    """discrete_sigma[t_to_idx(t)], so3_diffuser.py:288-297, 321-323, 398. t: numpy array (any float dtype)."""
    ds = discrete_sigma()
    idx = np.digitize(sigma_of_t(t), ds) - 1
    return ds[idx]


def quat_to_rotvec(quat, eps=1e-6):
    """framedipt/data/transforms.py:53-69."""
    flip = (quat[..., :1] < 0).to(quat.dtype)
    quat = (-1 * quat) * flip + (1 - flip) * quat
    angle = 2 * torch.atan2(torch.linalg.norm(quat[..., 1:], dim=-1), quat[..., 0])
    a2 = angle * angle
    small = 2 + a2 / 12 + 7 * a2 * a2 / 2880
    large = angle / torch.sin(angle / 2 + eps)
    sm = (angle <= 1e-3).to(quat.dtype)
    scale = small * sm + (1 - sm) * large
    return scale[..., None] * quat[..., 1:]


def rot_score(quats_t, quats_0, t):
    """SE3Diffuser.calc_rot_score (se3_diffuser.py:281-292) + SO3Diffuser.torch_score / igso3_expansion /
    score (so3_diffuser.py:373-402, 18-77, 122-191).  t: float32 tensor [B]. Returns float64 like the reference."""
    q0 = quats_0.clone()
    q0[..., 1:] *= -1
    q0 = q0 / (quats_0 ** 2).sum(-1, keepdim=True)
    v = quat_to_rotvec(quat_multiply(q0, quats_t))
    omega = torch.linalg.norm(v, dim=-1) + 1e-6
    sigma = torch.tensor(sigma_grid_value(t.cpu().numpy()))[:, None]  # float64 [B,1]
    lv = torch.arange(1000)[None, None]
    om = omega[..., None]
    eps = sigma[..., None]
    hi = torch.sin(om * (lv + 1 / 2))
    lo = torch.sin(om / 2)
    coef = (2 * lv + 1) * torch.exp(-lv * (lv + 1) * eps ** 2 / 2)
    expn = (coef * hi / lo).sum(-1)
    dhi = (lv + 1 / 2) * torch.cos(om * (lv + 1 / 2))
    dlo = 1 / 2 * torch.cos(om / 2)
    ds = (coef * (lo * dhi - hi * dlo) / lo ** 2).sum(-1)
    sc = ds / (expn + 1e-4)
    return sc[..., None] * v / omega[..., None]


def rot_score_conditioning(quats_t, quats_0, t):
    """Conditioning probe for tests.  The reference evaluates sin/cos((l+1/2) omega) of its 1000-term IGSO(3) series
    in float32 (arguments up to ~3000 rad), so each term carries ~1e-7 relative noise that depends on the libm used.
    Where the series cancels massively (omega >> sigma: the density is negligible; omega -> 0 or pi) the reference's
    own value is dominated by that noise and two equally valid libm implementations disagree.  Returns, per element,
    the expected libm-noise level of the score MAGNITUDE: 1.2e-7 * (sum|terms_df| / |f + 1e-4| + |df| sum|terms_f| / (f+1e-4)^2)
    and the magnitude itself."""
    q0 = quats_0.clone()
    q0[..., 1:] *= -1
    q0 = q0 / (quats_0 ** 2).sum(-1, keepdim=True)
    v = quat_to_rotvec(quat_multiply(q0, quats_t))
    omega = (torch.linalg.norm(v, dim=-1) + 1e-6).numpy().astype(np.float64)
    sig = sigma_grid_value(t.cpu().numpy())[:, None]
    l = np.arange(1000)
    lh = l + 0.5
    arg = omega[..., None] * lh
    coef = (2 * l + 1) * np.exp(-l * (l + 1) * sig[..., None] ** 2 / 2)
    hi, c = np.sin(arg), np.cos(arg)
    lo, dlo = np.sin(omega / 2)[..., None], 0.5 * np.cos(omega / 2)[..., None]
    tf = coef * hi / lo
    tdf_abs = coef * (np.abs(lo * lh * c) + np.abs(hi * dlo)) / lo ** 2  # magnitudes of the two cancelling products
    tdf = coef * (lo * lh * c - hi * dlo) / lo ** 2
    f, df = tf.sum(-1), tdf.sum(-1)
    den = np.abs(f + 1e-4)
    noise = 1.2e-7 * (tdf_abs.sum(-1) / den + np.abs(df) * np.abs(tf).sum(-1) / den ** 2)
    return noise, np.abs(df / (f + 1e-4))


def marginal_b_t(t):
    """r3_diffuser.py:87-96."""
    return t * MIN_B + 0.5 * (t ** 2) * (MAX_B - MIN_B)


def trans_score(trans_t, trans_0, t):
    """R3Diffuser.score with scale=True, use_torch=True (r3_diffuser.py:410-440); t: float32 [B,1,1]."""
    x_t, x_0 = trans_t * COORD_SCALE, trans_0 * COORD_SCALE
    mb = marginal_b_t(t)
    return -(x_t - torch.exp(-0.5 * mb) * x_0) / (1 - torch.exp(-mb))


# --------------------------------------------------------------------------------------------
# full forward  (score_network.py:218-275 + ipa_pytorch.py:509-572)
# --------------------------------------------------------------------------------------------
def score_network_forward(sd, feats, inpainting=True, input_aatype=True, dtype=torch.float32, taps=None):
    """ScoreNetwork.forward.  feats as in SURVEY §8 row A18.  Returns dict(rigids, rot_score, trans_score, psi)."""
    bb_mask = feats["res_mask"].to(torch.float32)
    fixed_mask = feats["fixed_mask"].to(torch.float32)
    aatype = preprocess_aatype(feats.get("aatype"), fixed_mask, inpainting, input_aatype)
    node0, edge0 = embedder(sd, feats["seq_idx"], feats["t"], fixed_mask, feats["sc_ca_t"].float(), aatype, dtype)
    node_mask = bb_mask.to(dtype)
    edge_mask = node_mask[..., None] * node_mask[..., None, :]
    z = edge0 * edge_mask[..., None]
    init_node = node0 * node_mask[..., None]
    if taps is not None:
        taps["node_embed0"], taps["edge_embed0"] = init_node.clone(), z.clone()
    diffuse_mask = ((1 - fixed_mask) * bb_mask).to(dtype)
    init = feats["rigids_t"].to(torch.float32).to(dtype)
    quats, trans = init[..., :4].clone(), init[..., 4:] * COORD_SCALE
    node = init_node
    for b in range(N_BLOCKS):
        p = "score_model.trunk."
        ipa_out = ipa(sd, f"{p}ipa_{b}", node, z, quats, trans, node_mask) * node_mask[..., None]
        if taps is not None:
            taps[f"ipa_{b}"] = ipa_out.clone()
        node = layer_norm(node + ipa_out, sd, f"{p}ipa_ln_{b}")
        x = torch.cat([node, linear(init_node, sd, f"{p}skip_embed_{b}")], -1)
        x = seq_transformer(sd, f"{p}seq_tfmr_{b}", x, node_mask)
        node = node + linear(x, sd, f"{p}post_tfmr_{b}")
        node = node_transition(sd, f"{p}node_transition_{b}", node) * node_mask[..., None]
        upd = linear(node * diffuse_mask[..., None], sd, f"{p}bb_update_{b}.linear")
        quats, trans = compose_q_update(quats, trans, upd, diffuse_mask[..., None])
        if taps is not None:
            taps[f"node_{b}"] = node.clone()
        if b < N_BLOCKS - 1:
            z = edge_transition(sd, f"{p}edge_transition_{b}", node, z) * edge_mask[..., None]
            if taps is not None:
                taps[f"edge_{b}"] = z.clone()
    rs = rot_score(init[..., :4], quats, feats["t"]) * node_mask[..., None]
    trans_out = trans / COORD_SCALE
    ts = trans_score(init[..., 4:], trans_out, feats["t"][:, None, None].to(dtype)) * node_mask[..., None]
    psi = torsion_angles(sd, "score_model.torsion_pred", node)
    gt_psi = feats["torsion_angles_sin_cos"][..., 2, :]
    dm = 1 - fixed_mask[..., None]
    psi = dm * psi + (1 - dm) * gt_psi
    return {"rigids": torch.cat([quats, trans_out], -1), "rot_score": rs, "trans_score": ts, "psi": psi}


# --------------------------------------------------------------------------------------------
# reverse step (numpy float64, like the reference)  se3_diffuser.py:346-401
# --------------------------------------------------------------------------------------------
def _rotvec_to_matrix(v):
    from scipy.spatial.transform import Rotation

    return Rotation.from_rotvec(v.reshape(-1, 3)).as_matrix().reshape(v.shape[:-1] + (3, 3))


def _matrix_to_rotvec(R):
    from scipy.spatial.transform import Rotation

    return Rotation.from_matrix(R.reshape(-1, 3, 3)).as_rotvec().reshape(R.shape[:-2] + (3,))


def so3_diffusion_coef(t):
    """so3_diffuser.py:308-319."""
    s = sigma_of_t(t)
    return np.sqrt(2 * (np.exp(MAX_SIGMA) - np.exp(MIN_SIGMA)) * s / np.exp(s))


def reverse_step(rigids_t, rot_score_t, trans_score_t, diffuse_mask, t, dt, z_rot, z_trans, center=True, noise_scale=1.0,
                 diffuse_rot=True, diffuse_trans=True):
    """SE3Diffuser.reverse with explicit noise (z_rot drawn before z_trans in the reference, §3.4).
    rigids_t: float32 tensor7 numpy [B,N,7]; scores numpy; z_*: N(0,1) numpy [B,N,3].
    Returns (rotmats float32 [B,N,3,3], trans float32 [B,N,3]) as _assemble_rigid stores them."""
    q = torch.tensor(rigids_t[..., :4], dtype=torch.float32)
    R_t = quat_to_rot(q).numpy()  # fp32 rot mats (rigid_utils.get_rot_mats)
    rot_t = _matrix_to_rotvec(R_t.astype(np.float64))
    trans_t = rigids_t[..., 4:]
    # SO3Diffuser.reverse, so3_diffuser.py:569-602
    g = so3_diffusion_coef(t)
    perturb = (g ** 2) * rot_score_t * dt + g * np.sqrt(dt) * (noise_scale * z_rot)
    rot_t_1 = _matrix_to_rotvec(np.einsum("...ij,...jk->...ik", _rotvec_to_matrix(rot_t), _rotvec_to_matrix(perturb)))
    # R3Diffuser.reverse, r3_diffuser.py:344-385
    x_t = trans_t * COORD_SCALE
    b_t = MIN_B + t * (MAX_B - MIN_B)
    g_t = np.sqrt(b_t)
    f_t = -0.5 * b_t * x_t
    perturb = (f_t - g_t ** 2 * trans_score_t) * dt + g_t * np.sqrt(dt) * (noise_scale * z_trans)
    perturb = perturb * diffuse_mask[..., None]
    x_t_1 = x_t - perturb
    if center:
        com = np.sum(x_t_1, axis=-2) / np.sum(diffuse_mask, axis=-1)[..., None]
        x_t_1 = x_t_1 - com[..., None, :]
    trans_t_1 = x_t_1 / COORD_SCALE
    if not diffuse_rot:  # se3_diffuser.py:373-385: an undiffused component is passed through unchanged
        rot_t_1 = rot_t
    if not diffuse_trans:
        trans_t_1 = trans_t
    dm = diffuse_mask[..., None]
    trans_t_1 = dm * trans_t_1 + (1 - dm) * trans_t
    rot_t_1 = dm * rot_t_1 + (1 - dm) * rot_t
    return _rotvec_to_matrix(rot_t_1).astype(np.float32), trans_t_1.astype(np.float32)


# --------------------------------------------------------------------------------------------
# backbone atoms  (all_atom.py:147-176, feats.py:165-228)
# --------------------------------------------------------------------------------------------
def compute_backbone(rotmats, trans, psi, aatype):
    """rotmats [B,N,3,3], trans [B,N,3] (Å), psi [B,N,2] = (sin, cos); aatype int [B,N] or None.
    Returns atom37 [B,N,37,3] float32 (slots N0 CA1 C2 CB3 O4; masked by any(!=0) like utils.py:415-438)."""
    rotmats = torch.as_tensor(rotmats, dtype=torch.float32)
    trans = torch.as_tensor(trans, dtype=torch.float32)
    psi = torch.as_tensor(psi, dtype=torch.float32)
    B, N = trans.shape[:2]
    aa = torch.zeros(B, N, dtype=torch.long) if aatype is None else torch.as_tensor(aatype).long().clone()
    aa[aa == 20] = 0
    pos = torch.tensor(_bc.IDEAL_BB_POS)[aa]  # [B,N,5,3] N,CA,C,O,CB
    frm = torch.tensor(_bc.PSI_DEFAULT_FRAME)[aa]  # [B,N,4,4]
    msk = torch.tensor(_bc.BB_ATOM_MASK)[aa]
    s, c = psi[..., 0], psi[..., 1]
    rx = torch.zeros(B, N, 3, 3)
    rx[..., 0, 0] = 1
    rx[..., 1, 1] = c
    rx[..., 1, 2] = -s
    rx[..., 2, 1] = s
    rx[..., 2, 2] = c
    Rpsi = frm[..., :3, :3] @ rx
    Rg = rotmats @ Rpsi
    tg = torch.einsum("bnij,bnj->bni", rotmats, frm[..., :3, 3]) + trans
    out = torch.zeros(B, N, 37, 3)
    bb = torch.einsum("bnij,bnaj->bnai", rotmats, pos) + trans[:, :, None, :]
    o = torch.einsum("bnij,bnj->bni", Rg, pos[:, :, 3]) + tg
    out[:, :, 0] = bb[:, :, 0] * msk[:, :, 0:1]
    out[:, :, 1] = bb[:, :, 1] * msk[:, :, 1:2]
    out[:, :, 2] = bb[:, :, 2] * msk[:, :, 2:3]
    out[:, :, 3] = bb[:, :, 4] * msk[:, :, 4:5]
    out[:, :, 4] = o * msk[:, :, 3:4]
    return out


# --------------------------------------------------------------------------------------------
# sampling loop  (experiments/utils.py:511-626, 292-412)
# --------------------------------------------------------------------------------------------
def inference_loop(sd, feats, num_t, min_t, noise, noise_scale=1.0, center=True, inpainting=True,
                   input_aatype=True, dtype=torch.float32, teacher=None, self_condition=True, diffuse_rot=True, diffuse_trans=True):
    """inference_fn with aux_traj=True, self_condition=True, embed_self_conditioning=True.
    noise: float64 [num_t-1, 2, B, N, 3] standard normals in the reference's draw order (rot, then trans).
    Returns dict(prot_traj [T,B,N,37,3], rigid_traj [T+1,B,N,7] (quats from scipy; sign may differ), rigid_0_traj)."""
    feats = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in feats.items()}
    B, N = feats["res_mask"].shape
    steps = np.linspace(min_t, 1.0, num_t)[::-1]
    dt = 1 / num_t
    ones = torch.ones(B)
    aatype = preprocess_aatype(feats.get("aatype"), feats["fixed_mask"], inpainting, input_aatype)
    diffuse_mask = ((1 - feats["fixed_mask"]) * feats["res_mask"]).numpy().astype(np.float64)
    rigid_traj = [feats["rigids_t"].numpy().copy()]
    prot, prot0 = [], []
    with torch.no_grad():
        if self_condition:  # experiments/utils.py:571-578: the pre-pass is the only thing `self_condition` switches
            feats["t"] = steps[0] * ones
            feats["sc_ca_t"] = score_network_forward(sd, feats, inpainting, input_aatype, dtype)["rigids"][..., 4:].float()
        for si, t in enumerate(steps):
            feats["t"] = t * ones
            out = score_network_forward(sd, feats, inpainting, input_aatype, dtype)
            rig_pred = out["rigids"].float()
            if t > min_t:
                feats["sc_ca_t"] = rig_pred[..., 4:]
                R1, T1 = reverse_step(feats["rigids_t"].float().numpy(), out["rot_score"].numpy().astype(np.float64),
                                      out["trans_score"].float().numpy(), diffuse_mask, t, dt, noise[si, 0], noise[si, 1],
                                      center=center, noise_scale=noise_scale, diffuse_rot=diffuse_rot, diffuse_trans=diffuse_trans)
            else:
                R1 = quat_to_rot(rig_pred[..., :4]).numpy()
                T1 = rig_pred[..., 4:].numpy()
            q1 = rot_to_quat_np(R1.astype(np.float64)).astype(np.float32)
            feats["rigids_t"] = torch.tensor(np.concatenate([q1, T1], -1))
            rigid_traj.append(feats["rigids_t"].numpy().copy())
            psi = out["psi"].float()
            prot0.append(compute_backbone(quat_to_rot(rig_pred[..., :4]), rig_pred[..., 4:], psi, aatype).numpy())
            prot.append(compute_backbone(R1, T1, psi, aatype).numpy())
    return {
        "prot_traj": np.flip(np.stack(prot), 0),
        "rigid_traj": np.flip(np.stack(rigid_traj), 0),
        "rigid_0_traj": np.flip(np.stack(prot0), 0),
        "psi_pred": psi.numpy()[None],
    }


# --------------------------------------------------------------------------------------------
# x_T sampling  (se3_diffuser.py:455-529, so3_diffuser.py:325-357, r3_diffuser.py:294-331)
# --------------------------------------------------------------------------------------------
def igso3_cdf_t1(num_omega=1000):
    """_cdf[t_to_idx(1.0)] computed lazily for the single sigma row used by sample_ref (so3_diffuser.py:247-262)."""
    omega = np.linspace(0, np.pi, num_omega + 1)[1:]
    sig = float(sigma_grid_value(np.array(1.0)))
    lv = np.arange(1000)[None]
    p = ((2 * lv + 1) * np.exp(-lv * (lv + 1) * sig ** 2 / 2) * np.sin(omega[:, None] * (lv + 0.5)) / np.sin(omega[:, None] / 2)).sum(-1)
    pdf = p * (1 - np.cos(omega)) / np.pi
    return omega, pdf.cumsum() / num_omega * np.pi


def sample_ref(n, gt_rotmats, gt_trans, diffuse_mask):
    """SE3Diffuser.sample_ref for one sample of n residues using the *global legacy numpy RNG* in the
    reference's draw order: randn(n,3), rand(n), normal([n_diffused,3])."""
    omega, cdf = igso3_cdf_t1()
    x = np.random.randn(n, 3)
    x /= np.linalg.norm(x, axis=-1, keepdims=True)
    ang = np.interp(np.random.rand(n), cdf, omega)
    rot_ref = x * ang[:, None]
    bm = diffuse_mask.astype(bool)
    x_ref = gt_trans.astype(np.float32) * COORD_SCALE
    loc = np.zeros_like(gt_trans[bm])
    inp = np.random.normal(loc=loc, scale=np.ones_like(loc))
    x_out = x_ref.copy()
    x_out[bm] = inp
    trans_ref = x_out / COORD_SCALE
    rot_imp = _matrix_to_rotvec(np.asarray(gt_rotmats, dtype=np.float32).astype(np.float64))
    dm = diffuse_mask[..., None]
    rot_ref = dm * rot_ref + (1 - dm) * rot_imp
    R = _rotvec_to_matrix(rot_ref).astype(np.float32)
    q = rot_to_quat_np(R.astype(np.float64)).astype(np.float32)
    return np.concatenate([q, trans_ref.astype(np.float32)], -1)
