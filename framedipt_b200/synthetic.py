"""Synthetic inputs for benchmarks and parity tests.

Real weights/structures are unavailable offline, so every workload is generated deterministically:
ground-truth frames = random rotations + a 3.8 Å random-walk CA trace; masks/chain layout per
BASELINE.json config; x_T from ``SE3Diffuser.sample_ref`` (host numpy, reference RNG order).
The feature dict follows the contract of the reference samplers (experiments/sampler.py:69-111,
267-354): res_mask/fixed_mask float64 [B,N], seq_idx/aatype int64, torsion_angles_sin_cos [B,N,7,2],
sc_ca_t zeros [B,N,3], rigids_t [B,N,7] (quat wxyz + trans Å), t.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
import torch

RESIDUE_GAP = 200  # framedipt/__init__.py:3


@dataclass
class Workload:
    name: str
    batch: int
    chains: tuple[int, ...]
    spans: tuple[tuple[int, int], ...]  # diffused [start, end) in global residue index; () = de novo (all diffused)
    num_t: int
    de_novo: bool = False
    noise_scale: float = 0.1
    min_t: float = 0.01

    @property
    def n_res(self) -> int:
        return int(sum(self.chains))


# BASELINE.json configs
WORKLOADS = {
    "cfg1_monomer64": Workload("cfg1_monomer64", 1, (64,), ((20, 32),), 50),
    "cfg2_tcr350": Workload("cfg2_tcr350", 8, (170, 180), ((90, 102), (260, 272)), 500),
    "cfg3_denovo256": Workload("cfg3_denovo256", 64, (256,), (), 500, de_novo=True),
    "cfg4_tcrpmhc800": Workload("cfg4_tcrpmhc800", 32, (210, 245, 9, 336), ((95, 107), (305, 317)), 500),
    "cfg5_sweep128": Workload("cfg5_sweep128", 128, (128,), (), 200, de_novo=True),
    "cfg5_sweep256": Workload("cfg5_sweep256", 128, (256,), (), 200, de_novo=True),
    "cfg5_sweep512": Workload("cfg5_sweep512", 128, (512,), (), 200, de_novo=True),
    "cfg5_sweep1024": Workload("cfg5_sweep1024", 128, (1024,), (), 200, de_novo=True),
}


def ground_truth(n_res: int, seed: int = 0):
    """Random rotations + 3.8 Å random-walk CA trace, centred. Returns (rotmats f32 [N,3,3], trans f32 [N,3])."""
    from scipy.spatial.transform import Rotation

    rs = np.random.RandomState(seed)
    R = Rotation.random(n_res, random_state=rs).as_matrix().astype(np.float32)
    steps = rs.normal(size=(n_res, 3))
    steps = 3.8 * steps / np.linalg.norm(steps, axis=-1, keepdims=True)
    ca = np.cumsum(steps, 0)
    # keep the walk compact (a folded chain, not a 3.8*sqrt(N) Å coil spreading without bound)
    ca = ca - ca.mean(0, keepdims=True)
    return R, ca.astype(np.float32)


def static_features(wl: Workload, seed: int = 0) -> dict[str, np.ndarray]:
    """Per-structure features shared by all B samples of a workload (host numpy)."""
    n = wl.n_res
    rs = np.random.RandomState(seed + 1)
    seq_idx = np.zeros(n, np.int64)
    chain_idx = np.zeros(n, np.int64)
    off = 0
    prev_len = 0
    for c, ln in enumerate(wl.chains):
        if wl.de_novo:
            seq_idx[off:off + ln] = np.arange(1, ln + 1)  # sampler.py:95 (1-based)
        else:
            seq_idx[off:off + ln] = prev_len + np.arange(ln)  # framedipt/data/utils.py:859-874
            prev_len += ln + RESIDUE_GAP
        chain_idx[off:off + ln] = c
        off += ln
    fixed = np.ones(n, np.float64)
    if wl.de_novo or not wl.spans:
        fixed[:] = 0.0
    for s, e in wl.spans:
        fixed[s:e] = 0.0
    ang = rs.uniform(-np.pi, np.pi, size=(n, 7))
    tors = np.stack([np.sin(ang), np.cos(ang)], -1).astype(np.float32)
    if wl.de_novo:
        tors[:] = 0.0
    R, ca = ground_truth(n, seed)
    return {
        "seq_idx": seq_idx, "chain_idx": chain_idx, "fixed_mask": fixed, "res_mask": np.ones(n, np.float64),
        "aatype": rs.randint(0, 20, size=n).astype(np.int64), "torsion_angles_sin_cos": tors,
        "gt_rotmats": R, "gt_trans": ca,
    }


def make_features(wl: Workload, diffuser, seed: int = 0, batch: int | None = None) -> dict[str, torch.Tensor]:
    """Batched feature dict on CPU. ``diffuser.sample_ref`` consumes the global numpy RNG (reference order), one
    call per sample exactly like the reference samplers (one `__getitem__` per sample, sampler.py:340-352)."""
    from .rigid import rotmats_to_rigid

    st = static_features(wl, seed)
    B = wl.batch if batch is None else batch
    n = wl.n_res
    dm = 1.0 - st["fixed_mask"]
    gt = rotmats_to_rigid(st["gt_rotmats"], st["gt_trans"])
    rig = []
    for _ in range(B):
        if wl.de_novo:
            r = diffuser.sample_ref(n_samples=n, as_tensor_7=True)["rigids_t"]
        else:
            r = diffuser.sample_ref(n_samples=n, impute=gt, diffuse_mask=dm, as_tensor_7=True)["rigids_t"]
        rig.append(torch.as_tensor(r).float())

    def rep(x):
        return torch.as_tensor(x)[None].repeat(B, *([1] * np.ndim(x)))

    feats = {
        "res_mask": rep(st["res_mask"]), "fixed_mask": rep(st["fixed_mask"]), "seq_idx": rep(st["seq_idx"]),
        "chain_idx": rep(st["chain_idx"]), "torsion_angles_sin_cos": rep(st["torsion_angles_sin_cos"]),
        "sc_ca_t": torch.zeros(B, n, 3), "rigids_t": torch.stack(rig), "t": torch.ones(B),
    }
    if not wl.de_novo:
        feats["aatype"] = rep(st["aatype"])
    return feats


def draw_noise(num_t: int, batch: int, n_res: int) -> np.ndarray:
    """Standard normals in the reference's draw order from the *global legacy numpy RNG*
    (so3_diffuser.py:591 then r3_diffuser.py:373, each [B,N,3], every step with t > min_t)."""
    return np.random.normal(size=(num_t - 1, 2, batch, n_res, 3))
