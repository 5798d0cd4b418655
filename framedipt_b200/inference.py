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
