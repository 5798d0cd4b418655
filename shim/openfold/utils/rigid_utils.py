"""openfold.utils.rigid_utils: the reference's own (pure torch) module when the checkout is present, else the small Rigid / Rotation
view types of framedipt_b200.rigid (the subset of the API the sampler path touches)."""
from framedipt_b200 import dropin as _d

_ref = _d.load_reference_module("openfold/utils/rigid_utils.py", "_framedipt_ref_rigid_utils")
if _ref is not None:
    globals().update({k: v for k, v in vars(_ref).items() if not k.startswith("__")})
else:
    from framedipt_b200.rigid import Rigid, Rotation  # noqa: F401
