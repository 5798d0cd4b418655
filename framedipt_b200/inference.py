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


class _TrajReader:
    """Streams fdpt_sample's trajectory slots to the host while later timesteps are still running: after every `chunk` steps the
    finished slots are copied device -> pinned staging on a side stream and scattered into the caller-owned numpy arrays (the
    backbone trajectories are expanded to the reference's [T,B,N,37,3] layout on the way; only 5 of the 37 atom slots are ever
    non-zero).  The reference materialises the same arrays with np.stack at the end (experiments/utils.py:610-626)."""

    def __init__(self, ctx, out: dict, T: int, chunk: int, keys: list[str], event_every: int | None = None):
        self.ctx, self.out, self.T, self.chunk, self.keys = ctx, out, T, chunk, keys
        self.event_every = event_every or chunk  # device-side progress events are recorded every this many timesteps
        self.side = torch.cuda.Stream(device=ctx.device)
        self.host = {}
        for k in keys:
            shp = tuple(out[k].shape)
            if k in ("prot_traj", "rigid_0_traj"):
                shp = shp[:-2] + (37, 3)
            self.host[k] = np.empty(shp, np.float32)
        self.stage = {k: torch.empty((min(chunk, T) + 1,) + tuple(out[k].shape[1:]), dtype=torch.float32).pin_memory() for k in keys}

    def _copy(self, key: str, lo: int, hi: int):
        """slots [lo, hi) of trajectory `key`"""
        n = hi - lo
        st = self.stage[key][:n]
        with torch.cuda.stream(self.side):
            st.copy_(self.out[key][lo:hi], non_blocking=True)
        self.side.synchronize()
        dst = self.host[key]
        if key in ("prot_traj", "rigid_0_traj"):
            dst[lo:hi, ..., :5, :] = st.numpy()
            dst[lo:hi, ..., 5:, :] = 0.0
        else:
            dst[lo:hi] = st.numpy()

    def run(self) -> dict[str, np.ndarray]:
        T, c = self.T, self.chunk
        # read-back groups: whole chunks while the trajectory is long, then single events: whatever is read after the last
        # timestep has finished is exposed time, so the final group is as small as the event granularity allows
        ev = self.event_every
        bounds = list(range(0, max(T - c, 0), c))
        b = bounds[-1] + c if bounds else 0
        while b < T:
            bounds.append(b)
            b += ev
        bounds.append(T)
        bounds = sorted(set(bounds))
        for s0, s1 in zip(bounds[:-1], bounds[1:]):
            self.ctx.wait_step(s1 - 1)
            lo, hi = T - s1, T - s0  # step s writes slot T-1-s (index 0 = final sample)
            for k in self.keys:
                if k == "rigid_traj":  # [T+1]: slot T is x_T (written before the loop), slots shift by none otherwise
                    self._copy(k, lo, hi + (1 if s0 == 0 else 0))
                else:
                    self._copy(k, lo, hi)
        return self.host


def inference_fn(model: ScoreNetwork, diffuser, data_init: dict, num_t: int, min_t: float, center: bool = True,
                 aux_traj: bool = False, self_condition: bool = True, noise_scale: float = 1.0,
                 embed_self_conditioning: bool = True, inpainting: bool = False, input_aatype: bool = False,
                 noise: np.ndarray | None = None, rng: str = "numpy", philox_seed: int | None = None,
                 readback_chunk: int = 25) -> dict[str, np.ndarray]:
    """experiments/utils.py:511-626 with the same arguments and return dict.  Extra keyword arguments (not in the reference):
    `noise` (pre-drawn [num_t-1, 2, B, N, 3] float64 normals), `rng` = "numpy" (default: the reference's legacy global numpy
    stream, bit-identical draws) or "philox" (throughput mode: normals drawn on the device, seed `philox_seed` or one drawn from
    numpy's global RNG), `readback_chunk` (timesteps per streamed trajectory read-back)."""
    if not isinstance(model, ScoreNetwork):
        raise TypeError("framedipt_b200.inference_fn drives framedipt_b200.ScoreNetwork (there is no eager fallback)")
    if rng not in ("numpy", "philox"):
        raise ValueError(f"rng should be 'numpy' or 'philox', got {rng}")
    feats = dict(data_init)  # shallow: nothing is mutated
    if feats["rigids_t"].ndim == 2:
        feats = {k: (v[None] if torch.is_tensor(v) and v.ndim >= 1 and k != "t" else v) for k, v in feats.items()}
    dev = feats["rigids_t"].device
    if dev.type != "cuda":
        raise runtime.FdptError("inference_fn needs the features on a CUDA device (no CPU fallback)")
    ctx = model.context(dev)
    # backbone residue types follow the CALL's flags (experiments/utils.py:549-555), the embedder's follow the model's
    from .score_network import preprocess_aatype

    aatype_bb = preprocess_aatype(feats.get("aatype"), torch.as_tensor(feats["fixed_mask"]), inpainting, input_aatype)
    pf = model.prepare(feats, dev, aatype_bb=aatype_bb)
    B, N = pf.B, pf.N
    steps, sched, t_emb_tab = build_schedule(diffuser, num_t, min_t, noise_scale)
    n_rev = int((sched[:, 7] == 0).sum())
    if n_rev != num_t - 1 or sched[-1, 7] != 1.0:
        raise ValueError("unexpected schedule: exactly the last step must satisfy t <= min_t")
    noise_d = None
    seed = 0
    if noise is not None or rng == "numpy":
        if noise is None:
            noise = draw_noise(diffuser, n_rev, B, N)
        noise_d = torch.as_tensor(np.ascontiguousarray(noise, np.float64)).to(dev) if n_rev > 0 else torch.zeros(1, 2, B, N, 3, dtype=torch.float64, device=dev)
    else:
        seed = int(np.random.randint(0, 2 ** 31 - 1)) if philox_seed is None else int(philox_seed)
    flags = (1 if (embed_self_conditioning and self_condition) else 0) | (0 if embed_self_conditioning else 2)
    chunk = max(1, min(int(readback_chunk), num_t))
    ev = 5 if chunk % 5 == 0 else chunk  # progress events every 5 timesteps: the last read-back group covers 5 steps, not `chunk`
    out = ctx.sample(pf, sched, t_emb_tab, noise_d, self_condition=flags, center=center, diffuse_rot=diffuser._diffuse_rot,
                     diffuse_trans=diffuser._diffuse_trans, final_only=False, philox_seed=seed, progress_chunk=ev)
    keys = ["prot_traj"] + (["rigid_traj", "trans_traj", "rigid_0_traj"] if aux_traj else [])
    host = _TrajReader(ctx, out, num_t, chunk, keys, event_every=ev).run()
    torch.cuda.current_stream(dev).synchronize()
    ret = {"prot_traj": host["prot_traj"]}
    if aux_traj:
        ret["rigid_traj"] = host["rigid_traj"]
        ret["trans_traj"] = host["trans_traj"]
        ret["psi_pred"] = out["psi_pred"].cpu().numpy()[None]
        ret["rigid_0_traj"] = host["rigid_0_traj"]
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
