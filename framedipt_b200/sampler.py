"""Samplers with the reference's `(key, sample_i, feats)` tuple contract (experiments/sampler.py:21-135, 137-354) and the
batched, device-side construction of x_T.

* ``UnconditionalSampler(cfg, diffuser, device)`` — de-novo lengths sweep, exactly the reference's class (host-only inputs).
* ``SyntheticConditionalSampler(data_conf, diffuser, device)`` — stands in for ``ConditionalSampler`` / ``TCRSampler`` when no mmCIF
  data / Bio / ANARCI is available (benchmarks, tests): structures come from ``synthetic.static_features`` instead of processed
  PDB entries; the feature dict has the reference's keys, dtypes and leading batch dimension of 1.
* ``sample_ref_batch`` / ``batch_features`` — B samples of ONE structure built on the GPU in one call (the reference makes B host
  ``sample_ref`` calls): x_T by ``fdpt_sample_ref`` (parity mode: the legacy numpy stream drawn on the host in the reference's order;
  throughput mode: device Philox), static features replicated on the device.
* ``pad_feats`` / ``pad_rigid`` — framedipt/data/utils.py:311-339, for batching different structures.
"""
from __future__ import annotations

import numpy as np
import torch

from . import synthetic
from .rigid import Rigid, rotmats_to_rigid

UNPADDED_FEATS = ["t", "rot_score_scaling", "trans_score_scaling", "t_seq", "t_struct"]  # framedipt/data/utils.py:41-43
RIGID_FEATS = ["rigids_0", "rigids_t"]
PAIR_FEATS = ["rel_rots"]


# ------------------------------------------------------------------------------------------------------------------
# padding (framedipt/data/utils.py:311-380)
# ------------------------------------------------------------------------------------------------------------------
def pad(x, max_len: int, pad_idx: int = 0, use_torch: bool = False, reverse: bool = False):
    pad_amt = max_len - x.shape[pad_idx]
    if pad_amt < 0:
        raise ValueError(f"Invalid pad amount {pad_amt}")
    widths = [(0, 0)] * x.ndim
    widths[pad_idx] = (pad_amt, 0) if reverse else (0, pad_amt)
    if use_torch:
        return torch.nn.functional.pad(x, sum(widths[::-1], ()))
    return np.pad(x, widths)


def pad_rigid(rigid: torch.Tensor, max_len: int) -> torch.Tensor:
    ident = Rigid.identity((max_len - rigid.shape[0],), device=rigid.device).to_tensor_7().to(rigid.dtype)
    return torch.cat([rigid, ident], dim=0)


def pad_feats(raw_feats: dict, max_len: int, use_torch: bool = False) -> dict:
    out = {k: pad(v, max_len, use_torch=use_torch) for k, v in raw_feats.items() if k not in UNPADDED_FEATS + RIGID_FEATS}
    for k in PAIR_FEATS:
        if k in out:
            out[k] = pad(out[k], max_len, pad_idx=1)
    for k in UNPADDED_FEATS:
        if k in raw_feats:
            out[k] = raw_feats[k]
    for k in RIGID_FEATS:
        if k in raw_feats:
            out[k] = pad_rigid(raw_feats[k], max_len)
    return out


# ------------------------------------------------------------------------------------------------------------------
# x_T on the device
# ------------------------------------------------------------------------------------------------------------------
def sample_ref_batch(ctx, diffuser, B: int, n_res: int, gt_rotmats=None, gt_trans=None, diffuse_mask=None, rng: str = "numpy",
                     philox_seed: int = 0) -> torch.Tensor:
    """[B, n_res, 7] x_T of B samples of one structure (SE3Diffuser.sample_ref semantics, se3_diffuser.py:455-529).
    rng="numpy": the draws come from the legacy global numpy RNG in the reference's order, per sample randn(n,3), rand(n),
    normal([n_diffused,3]) — so the result equals B sequential reference calls; rng="philox": drawn on the device."""
    dev = ctx.device
    so3 = diffuser._so3_diffuser
    cdf = torch.as_tensor(so3.cdf_row(int(so3.t_to_idx(1.0)))).to(dev)
    omg = torch.as_tensor(so3.discrete_omega).to(dev)
    impute = dm_d = None
    if gt_rotmats is not None:
        impute = rotmats_to_rigid(gt_rotmats, gt_trans).to_tensor_7().to(dev, torch.float32).contiguous()
        dm = np.ones(n_res) if diffuse_mask is None else np.asarray(diffuse_mask, np.float64)
        dm_d = torch.as_tensor(dm, dtype=torch.float32).to(dev)
    elif diffuse_mask is not None:
        raise ValueError("Must provide imputation values for unmasked regions!")
    draws = None
    if rng == "numpy":
        bm = np.ones(n_res, bool) if diffuse_mask is None else np.asarray(diffuse_mask).astype(bool)
        host = np.zeros((B, 7 * n_res))
        for b in range(B):
            if diffuser._diffuse_rot:
                host[b, :3 * n_res] = np.random.randn(n_res, 3).reshape(-1)
                host[b, 3 * n_res:4 * n_res] = np.random.rand(n_res)
            else:
                host[b, :3 * n_res] = 1.0
            if diffuser._diffuse_trans:
                z = np.zeros((n_res, 3))
                z[bm] = np.random.normal(loc=np.zeros((int(bm.sum()), 3)), scale=np.ones((int(bm.sum()), 3)))
                host[b, 4 * n_res:] = z.reshape(-1)
        draws = torch.as_tensor(host).to(dev)
    elif rng != "philox":
        raise ValueError(f"rng should be 'numpy' or 'philox', got {rng}")
    return ctx.sample_ref(B, n_res, impute, dm_d, cdf, omg, draws, philox_seed, diffuser._diffuse_rot, diffuser._diffuse_trans)


def batch_features(ctx, diffuser, wl: synthetic.Workload, seed: int = 0, batch: int | None = None, rng: str = "numpy",
                   philox_seed: int = 0) -> dict[str, torch.Tensor]:
    """The feature dict of `synthetic.make_features`, built on the device: same keys / dtypes, x_T from `sample_ref_batch`."""
    st = synthetic.static_features(wl, seed)
    B = wl.batch if batch is None else batch
    n = wl.n_res
    dev = ctx.device
    if wl.de_novo:
        rig = sample_ref_batch(ctx, diffuser, B, n, rng=rng, philox_seed=philox_seed)
    else:
        rig = sample_ref_batch(ctx, diffuser, B, n, st["gt_rotmats"], st["gt_trans"], 1.0 - st["fixed_mask"], rng=rng, philox_seed=philox_seed)

    def rep(x):
        x = torch.as_tensor(x).to(dev)
        return x[None].expand(B, *x.shape).contiguous()

    feats = {"res_mask": rep(st["res_mask"]), "fixed_mask": rep(st["fixed_mask"]), "seq_idx": rep(st["seq_idx"]), "chain_idx": rep(st["chain_idx"]),
             "torsion_angles_sin_cos": rep(st["torsion_angles_sin_cos"]), "sc_ca_t": torch.zeros(B, n, 3, device=dev), "rigids_t": rig,
             "t": torch.ones(B, device=dev)}
    if not wl.de_novo:
        feats["aatype"] = rep(st["aatype"])
    return feats


# ------------------------------------------------------------------------------------------------------------------
# datasets with the reference's item contract
# ------------------------------------------------------------------------------------------------------------------
def _tensorise(d: dict, device) -> dict:
    return {k: (v if torch.is_tensor(v) else torch.tensor(v))[None].to(device) for k, v in d.items()}


class UnconditionalSampler(torch.utils.data.Dataset):
    """experiments/sampler.py:21-135: de-novo design; items are (sample_length, sample_i, feats) with a leading batch dim of 1.
    cfg: min_length, max_length, length_step, samples_per_length."""

    def __init__(self, cfg, diffuser, device) -> None:
        self._cfg, self._diffuser, self.device = cfg, diffuser, device
        self.all_sampling_lengths = np.repeat(np.arange(cfg.min_length, cfg.max_length + 1, cfg.length_step), cfg.samples_per_length)

    def sample(self, sample_length: int) -> dict[str, torch.Tensor]:
        ref = self._diffuser.sample_ref(n_samples=sample_length, as_tensor_7=True)
        init = {"res_mask": np.ones(sample_length), "seq_idx": torch.arange(1, sample_length + 1), "fixed_mask": np.zeros(sample_length),
                "torsion_angles_sin_cos": np.zeros((sample_length, 7, 2)), "sc_ca_t": np.zeros((sample_length, 3)), **ref}
        return _tensorise(init, self.device)

    def __len__(self) -> int:
        return len(self.all_sampling_lengths)

    def __getitem__(self, item: int):
        if item >= len(self):
            raise IndexError(item)
        n = int(self.all_sampling_lengths[item])
        return n, item % self._cfg.samples_per_length, self.sample(n)


class SyntheticConditionalSampler(torch.utils.data.Dataset):
    """Inpainting items `(pdb_name, sample_idx, feats)` like ConditionalSampler.__getitem__ (experiments/sampler.py:267-354), from
    synthetic structures.  data_conf: `workloads` (names in synthetic.WORKLOADS or Workload objects), `samples` (per structure),
    `seed`.  Features carry the reference's keys: aatype, seq_idx, chain_idx, res_mask, fixed_mask, torsion_angles_sin_cos, rigids_0,
    sc_ca_t, rigids_t, t — one `sample_ref` call per item, legacy numpy RNG, exactly the reference's consumption."""

    def __init__(self, data_conf, diffuser, device) -> None:
        self._data_conf, self._diffuser, self.device = data_conf, diffuser, device
        wls = getattr(data_conf, "workloads", None) or data_conf["workloads"]
        self.workloads = [synthetic.WORKLOADS[w] if isinstance(w, str) else w for w in wls]
        self.samples = int(getattr(data_conf, "samples", 1))
        self.seed = int(getattr(data_conf, "seed", 0))

    @property
    def diffuser(self):
        return self._diffuser

    def __len__(self) -> int:
        return len(self.workloads) * self.samples

    def __getitem__(self, idx: int):
        if idx >= len(self):
            raise IndexError(idx)
        example_idx, sample_idx = divmod(idx, self.samples)
        wl = self.workloads[example_idx]
        st = synthetic.static_features(wl, self.seed + example_idx)
        gt = rotmats_to_rigid(st["gt_rotmats"], st["gt_trans"])
        diffused = 1.0 - st["fixed_mask"]
        if diffused.sum() < 1:
            raise ValueError("Must be diffused")
        feats = {"aatype": st["aatype"], "seq_idx": st["seq_idx"], "chain_idx": st["chain_idx"], "res_mask": st["res_mask"],
                 "fixed_mask": st["fixed_mask"], "torsion_angles_sin_cos": st["torsion_angles_sin_cos"], "rigids_0": gt.to_tensor_7(),
                 "sc_ca_t": torch.zeros_like(gt.get_trans())}
        feats.update(self._diffuser.sample_ref(n_samples=wl.n_res, chain_index=st["chain_idx"], impute=gt, diffuse_mask=diffused,
                                               as_tensor_7=True))
        feats["t"] = 1.0
        final = {k: (v if torch.is_tensor(v) else torch.tensor(v)) for k, v in feats.items()}
        final = pad_feats(final, wl.n_res, use_torch=True)
        return wl.name, sample_idx, {k: v[None].to(self.device) for k, v in final.items()}
