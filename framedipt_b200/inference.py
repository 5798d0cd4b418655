"""`inference_fn` — the reverse-diffusion sampling loop with the reference's signature
(experiments/utils.py:511-626), executed as ONE enqueue of libfdpt's `fdpt_sample`: the state stays on the
GPU for the whole trajectory (the reference round-trips to the host every step).

Noise: the reference draws `np.random.normal` from the legacy *global* numpy RNG, rot then trans, shape
[B,N,3] each, at every step with t > min_t (so3_diffuser.py:591, r3_diffuser.py:373).  The same stream is
drawn here (one vectorised call, bit-identical to the sequential calls) and uploaded.
"""
from __future__ import annotations

import numpy as np
import torch

from . import runtime
from .score_network import ScoreNetwork


def build_schedule(diffuser, num_t: int, min_t: float, noise_scale: float):
    """Per-step scalar table + timestep-embedding table (host; reference expressions)."""
    steps = np.linspace(min_t, 1.0, num_t)[::-1]
    dt = 1 / num_t
    sched = np.zeros((num_t, runtime.SCHED_COLS), np.float64)
    for i, t in enumerate(steps):
        sched[i] = diffuser.step_scalars(float(t), dt, noise_scale)
        sched[i, 7] = 0.0 if t > min_t else 1.0  # experiments/utils.py:352 `if t > min_t`
    t32 = torch.tensor(steps.astype(np.float32))
    return steps, sched, runtime.timestep_embedding(t32)


def draw_noise(diffuser, n_rev: int, B: int, N: int) -> np.ndarray:
    """[n_rev, 2, B, N, 3] float64 standard normals in the reference's draw order."""
    if diffuser._diffuse_rot and diffuser._diffuse_trans:
        return np.random.normal(size=(n_rev, 2, B, N, 3))
    z = np.zeros((n_rev, 2, B, N, 3))
    for s in range(n_rev):
        if diffuser._diffuse_rot:
            z[s, 0] = np.random.normal(size=(B, N, 3))
        if diffuser._diffuse_trans:
            z[s, 1] = np.random.normal(size=(B, N, 3))
    return z


def _expand37(bb5: np.ndarray) -> np.ndarray:
    out = np.zeros(bb5.shape[:-2] + (37, 3), np.float32)
    out[..., :5, :] = bb5
    return out


def inference_fn(model: ScoreNetwork, diffuser, data_init: dict, num_t: int, min_t: float, center: bool = True,
                 aux_traj: bool = False, self_condition: bool = True, noise_scale: float = 1.0,
                 embed_self_conditioning: bool = True, inpainting: bool = False, input_aatype: bool = False,
                 noise: np.ndarray | None = None) -> dict[str, np.ndarray]:
    if not isinstance(model, ScoreNetwork):
        raise TypeError("framedipt_b200.inference_fn drives framedipt_b200.ScoreNetwork (there is no eager fallback)")
    if not embed_self_conditioning:
        raise NotImplementedError("embed_self_conditioning=False is not supported by the CUDA path")
    feats = dict(data_init)  # shallow: nothing is mutated
    if feats["rigids_t"].ndim == 2:
        feats = {k: (v[None] if torch.is_tensor(v) and v.ndim >= 1 and k != "t" else v) for k, v in feats.items()}
    dev = feats["rigids_t"].device
    if dev.type != "cuda":
        raise runtime.FdptError("inference_fn needs the features on a CUDA device (no CPU fallback)")
    ctx = model.context(dev)
    pf = model.prepare(feats, dev)
    B, N = pf.B, pf.N
    steps, sched, t_emb_tab = build_schedule(diffuser, num_t, min_t, noise_scale)
    n_rev = int((sched[:, 7] == 0).sum())
    if n_rev != num_t - 1 or sched[-1, 7] != 1.0:
        raise ValueError("unexpected schedule: exactly the last step must satisfy t <= min_t")
    if noise is None:
        noise = draw_noise(diffuser, n_rev, B, N)
    noise_d = torch.as_tensor(np.ascontiguousarray(noise, np.float64)).to(dev) if n_rev > 0 else None
    out = ctx.sample(pf, sched, t_emb_tab, noise_d, self_condition=self_condition, center=center,
                     diffuse_rot=diffuser._diffuse_rot, diffuse_trans=diffuser._diffuse_trans, final_only=False)
    torch.cuda.current_stream(dev).synchronize()
    ret = {"prot_traj": _expand37(out["prot_traj"].cpu().numpy())}
    if aux_traj:
        ret["rigid_traj"] = out["rigid_traj"].cpu().numpy()
        ret["trans_traj"] = out["trans_traj"].cpu().numpy()
        ret["psi_pred"] = out["psi_pred"].cpu().numpy()[None]
        ret["rigid_0_traj"] = _expand37(out["rigid_0_traj"].cpu().numpy())
    return ret


# ---- EigenFold confidence score (experiments/utils.py:251-289, 752-869) -------------------------------------------------------
def one_step_inference_score(model: ScoreNetwork, diffuser, sample_feats: dict, t: float, t_placeholder: torch.Tensor,
                             self_condition: bool = True):
    """Translation and rotation scores of one network evaluation at time t (with the self-conditioning pre-pass)."""
    sample_feats["t"] = t * t_placeholder
    rot_scaling, trans_scaling = diffuser.score_scaling(t)
    sample_feats["rot_score_scaling"] = rot_scaling * t_placeholder
    sample_feats["trans_score_scaling"] = trans_scaling * t_placeholder
    if self_condition:
        sample_feats["sc_ca_t"] = model(sample_feats)["rigids"][..., 4:]
    out = model(sample_feats)
    return out["trans_score"], out["rot_score"]


def logp_confidence_score(model: ScoreNetwork, diffuser, rigids_t, sample_feats: dict, diffuse_mask: np.ndarray, num_t: int,
                          min_t: float, device, self_condition: bool):
    """log p(sample) estimated along one forward-noising path: sum over steps of log p(x(t-1)|x(t)) - log q(x(t)|x(t-1)), plus the
    prior log-density at t=1.  Same signature, RNG consumption (legacy global numpy stream: translation then rotation noise per
    step) and in-place updates of `sample_feats` as the reference; the two network evaluations per step run on the GPU through
    `ScoreNetwork`, the [N, 3]-sized transition densities are host numpy exactly as in the reference.
    Returns (log_prob, running log_prob after every step)."""
    from .se3_diffuser import _rotvec_of, gaussian_log_prob

    if not isinstance(model, ScoreNetwork):
        raise TypeError("framedipt_b200.logp_confidence_score drives framedipt_b200.ScoreNetwork (there is no eager fallback)")
    diffuse_mask = np.asarray(diffuse_mask)
    forward_steps = np.linspace(min_t, 1.0, num_t)[:-1]
    t_placeholder = torch.ones((1,)).to(device)
    dt = 1 / num_t
    log_probs, log_prob = [], 0.0
    for i, t_1 in enumerate(forward_steps):
        prev = rigids_t
        rigids_t = diffuser.forward(rigids_t_1=prev, t_1=t_1, diffuse_mask=diffuse_mask, dt=dt)
        sample_feats["rigids_t"] = rigids_t.to_tensor_7().to(device)
        if sample_feats["rigids_t"].ndim == 2 and sample_feats["res_mask"].ndim == 2:
            sample_feats["rigids_t"] = sample_feats["rigids_t"][None]
        t = 1.0 if i == len(forward_steps) - 1 else forward_steps[i + 1]
        trans_score, rot_score = one_step_inference_score(model, diffuser, sample_feats, t, t_placeholder, self_condition)
        trans_score = trans_score.squeeze().cpu().numpy()
        rot_score = rot_score.squeeze().cpu().numpy()
        log_prob += diffuser.log_prob_backward(rigids_t=rigids_t, rigids_t_1=prev, trans_score_t=trans_score, rot_score_t=rot_score,
                                               dt=dt, t=t, diffuse_mask=diffuse_mask)
        log_prob -= diffuser.log_prob_forward(rigids_t=rigids_t, rigids_t_1=prev, dt=dt, t_1=t_1, diffuse_mask=diffuse_mask)
        log_probs.append(log_prob)
    # prior at t = 1: standard normal on the scaled translations, uniform rotations
    trans, _ = _rotvec_of(rigids_t)
    trans = diffuser._r3_diffuser._scale(trans)
    log_prob += gaussian_log_prob(np.zeros_like(trans), np.ones_like(trans), trans, diffuse_mask)
    log_prob += np.log(1 / np.pi ** 2) * diffuse_mask.sum().item()
    log_probs.append(log_prob)
    return log_prob, log_probs
