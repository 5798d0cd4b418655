"""framedipt_b200 — B200-native SE(3) backbone-frame diffusion sampler hot path (drop-in for FrameDiPT's)."""
from .params import ModelDims, param_specs, synthetic_state_dict  # noqa: F401
from .rigid import Rigid, Rotation  # noqa: F401
from .se3_diffuser import SE3Diffuser  # noqa: F401

__all__ = ["ModelDims", "param_specs", "synthetic_state_dict", "Rigid", "Rotation", "SE3Diffuser", "ScoreNetwork", "inference_fn", "logp_confidence_score"]


def __getattr__(name):  # lazy: these need the CUDA library
    if name == "ScoreNetwork":
        from .score_network import ScoreNetwork

        return ScoreNetwork
    if name == "inference_fn":
        from .inference import inference_fn

        return inference_fn
    if name == "logp_confidence_score":
        from .inference import logp_confidence_score

        return logp_confidence_score
    raise AttributeError(name)
