"""experiments.utils: every helper of the reference's own module (when the checkout is importable), with the two entry points of the
sampler hot path replaced by the B200 path: `inference_fn` (experiments/utils.py:511-626) and `logp_confidence_score` (752-869)."""
from framedipt_b200 import dropin as _d

_ref = _d.load_reference_module("experiments/utils.py", "_framedipt_ref_experiments_utils")
if _ref is not None:
    globals().update({k: v for k, v in vars(_ref).items() if not k.startswith("__")})

from framedipt_b200.inference import inference_fn, logp_confidence_score  # noqa: E402,F401

if _ref is None:
    import torch as _torch

    def get_atom_positions_from_rigids(rigids, psi_torsions, aatype=None):
        """experiments/utils.py:415-438 on the device: [.., N, 37, 3] numpy (slots 0..4 = N, CA, C, CB, O)."""
        import numpy as np

        from framedipt_b200 import runtime

        r7 = rigids.to_tensor_7()
        dev = r7.device if r7.is_cuda else _torch.device("cuda", _torch.cuda.current_device())
        ctx = runtime.default_context(dev.index)
        sq = r7.dim() == 2
        r7 = (r7[None] if sq else r7).to(dev, _torch.float32).contiguous()
        psi = _torch.as_tensor(psi_torsions).to(dev, _torch.float32)
        psi = (psi[None] if sq else psi).contiguous()
        aa = None if aatype is None else (_torch.as_tensor(aatype)[None] if sq else _torch.as_tensor(aatype)).to(dev, _torch.int32).contiguous()
        bb = ctx.backbone(r7, psi, aa).cpu().numpy()
        out = np.zeros(bb.shape[:-2] + (37, 3), np.float32)
        out[..., :5, :] = bb
        return out[0] if sq else out
