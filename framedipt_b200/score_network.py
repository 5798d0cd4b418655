"""`ScoreNetwork` with the reference's call surface (framedipt/model/score_network.py:200-275), executed by libfdpt.so.

    model = ScoreNetwork(model_conf, diffuser, inpainting=True)
    model.load_state_dict(ckpt["model"]); model = model.to("cuda"); model.eval()
    out = model(input_feats)   # dict: psi, rot_score (float64), trans_score, rigids, atom37, atom14

Parameters are held as ordinary (frozen) ``nn.Parameter``s under the reference's state_dict names, so
``load_state_dict`` / ``state_dict`` / ``.to`` behave like the reference module; they are pushed into the CUDA
context lazily on the first call after a change.  Inference only (no autograd through the kernels).
"""
from __future__ import annotations

import numpy as np
import torch
from torch import nn

from . import runtime
from .params import ModelDims, synthetic_state_dict
from .se3_diffuser import _get


def preprocess_aatype(aatype, fixed_mask, inpainting: bool, input_aatype: bool):
    """framedipt/data/utils.py:565-610."""
    if aatype is None or (not inpainting and not input_aatype):
        return None
    aatype = aatype.type(torch.int64)
    if not input_aatype:
        fixed_mask = torch.as_tensor(fixed_mask).to(aatype.device)
        aatype = torch.where(fixed_mask.bool(), aatype, torch.full(aatype.shape, 20, dtype=torch.int64, device=aatype.device))
    return aatype


def dims_from_conf(model_conf) -> ModelDims:
    if model_conf is None:
        return ModelDims()
    ipa, emb = model_conf.ipa, model_conf.embed
    return ModelDims(c_s=int(ipa.c_s), c_z=int(ipa.c_z), c_hidden=int(ipa.c_hidden), c_skip=int(ipa.c_skip), no_heads=int(ipa.no_heads),
                     no_qk_points=int(ipa.no_qk_points), no_v_points=int(ipa.no_v_points), num_blocks=int(ipa.num_blocks),
                     index_embed_size=int(emb.index_embed_size), num_bins=int(emb.num_bins), min_bin=float(emb.min_bin),
                     max_bin=float(emb.max_bin), seq_tfmr_num_heads=int(ipa.seq_tfmr_num_heads),
                     seq_tfmr_num_layers=int(ipa.seq_tfmr_num_layers), coordinate_scaling=float(ipa.coordinate_scaling),
                     embed_self_conditioning=bool(_get(emb, "embed_self_conditioning", True)))


def _register(root: nn.Module, key: str, value: torch.Tensor):
    parts = key.split(".")
    mod = root
    for p in parts[:-1]:
        if p not in mod._modules:
            mod.add_module(p, nn.Module())
        mod = mod._modules[p]
    mod.register_parameter(parts[-1], nn.Parameter(value, requires_grad=False))


class ScoreNetwork(nn.Module):
    def __init__(self, model_conf, diffuser, inpainting: bool = False) -> None:
        super().__init__()
        self._model_conf = model_conf
        self.diffuser = diffuser
        self.inpainting = inpainting
        self._input_aatype = bool(_get(model_conf, "input_aatype", False)) if model_conf is not None else True
        self._with_aatype = bool(inpainting or self._input_aatype)
        self._dims = dims_from_conf(model_conf)
        for k, v in synthetic_state_dict(0, self._dims, self._with_aatype).items():
            _register(self, k, v)
        self._ctx: runtime.Context | None = None
        self._dirty = True

    # ---- nn.Module protocol used by the reference's callers (experiments/inference.py:149-161)
    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        res = super().load_state_dict(state_dict, strict=strict, **kw)
        self._dirty = True
        return res

    def _apply(self, fn, *a, **k):
        r = super()._apply(fn, *a, **k)
        self._dirty = True
        return r

    def context(self, device: torch.device) -> runtime.Context:
        idx = device.index if device.index is not None else torch.cuda.current_device()
        if self._ctx is None or self._ctx.device.index != idx:
            r3 = self.diffuser._r3_diffuser if self.diffuser is not None else None
            self._ctx = runtime.Context(self._dims, self._with_aatype, idx, r3.min_b if r3 else 0.1, r3.max_b if r3 else 20.0,
                                        r3.coordinate_scaling if r3 else 0.1)
            so3 = self.diffuser._so3_diffuser if self.diffuser is not None else None
            if so3 is not None and so3.use_cached_score:  # so3_diffuser.py:389-396: table look-up instead of the series
                self._ctx.set_score_table(so3.score_norms, so3.discrete_omega[:-1])
            self._dirty = True
        if self._dirty:
            self._ctx.load_state_dict(dict(self.state_dict()), strict=False)
            self._dirty = False
        return self._ctx

    def prepare(self, input_feats: dict, device: torch.device, aatype_bb=False) -> runtime.PreparedFeats:
        """aatype_bb: residue types for the trajectory's backbone atoms when they are decided by the caller's own flags
        (inference_fn, experiments/utils.py:549-555); False = follow the model's."""
        fixed_mask = torch.as_tensor(input_feats["fixed_mask"]).type(torch.float32)
        aatype = preprocess_aatype(input_feats.get("aatype"), fixed_mask, self.inpainting, self._input_aatype)
        return runtime.PreparedFeats(input_feats, device, self._dims, self._with_aatype, aatype, aatype_bb)

    @torch.no_grad()
    def forward(self, input_feats: dict[str, torch.Tensor]) -> dict[str, torch.Tensor]:
        dev = input_feats["rigids_t"].device
        if dev.type != "cuda":
            raise runtime.FdptError("framedipt_b200.ScoreNetwork runs on CUDA only (no CPU fallback); move the features to the GPU")
        ctx = self.context(dev)
        pf = self.prepare(input_feats, dev)
        t = input_feats["t"]
        t_np = torch.as_tensor(t).detach().to("cpu", torch.float32).numpy()
        so3 = self.diffuser._so3_diffuser
        out = ctx.forward(pf, t, np.asarray(so3.grid_sigma(t_np), np.float64).reshape(-1), sigma_idx=np.asarray(so3.t_to_idx(t_np)).reshape(-1))
        bb = out.pop("atom37_bb")
        B, N = pf.B, pf.N
        atom37 = torch.zeros(B, N, 37, 3, device=dev)
        atom37[:, :, :5] = bb
        atom14 = torch.zeros(B, N, 14, 3, device=dev)
        atom14[:, :, :3] = bb[:, :, :3]
        atom14[:, :, 3] = bb[:, :, 4]
        atom14[:, :, 4] = bb[:, :, 3]
        return {"psi": out["psi"], "rot_score": out["rot_score"], "trans_score": out["trans_score"], "rigids": out["rigids"],
                "atom37": atom37, "atom14": atom14}
