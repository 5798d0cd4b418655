"""framedipt.diffusion.se3_diffuser of the reference (framedipt/diffusion/se3_diffuser.py), served by the B200 path."""
from framedipt_b200.se3_diffuser import SE3Diffuser  # noqa: F401
from framedipt_b200.se3_diffuser import _rigid_of as _assemble_rigid  # noqa: F401  (se3_diffuser.py:26-36)
from framedipt_b200.se3_diffuser import _rotvec_of as _extract_trans_rots  # noqa: F401  (se3_diffuser.py:16-23)
