"""experiments.sampler: the reference's own samplers when their dependencies (Bio, ANARCI, processed mmCIF data) import; the de-novo
sampler (pure host code, same class) and the synthetic inpainting sampler of framedipt_b200.sampler otherwise / in addition."""
from framedipt_b200 import dropin as _d

_ref = _d.load_reference_module("experiments/sampler.py", "_framedipt_ref_experiments_sampler")
if _ref is not None:
    globals().update({k: v for k, v in vars(_ref).items() if not k.startswith("__")})
else:
    from framedipt_b200.sampler import UnconditionalSampler  # noqa: F401

from framedipt_b200.sampler import SyntheticConditionalSampler, batch_features, sample_ref_batch  # noqa: E402,F401
