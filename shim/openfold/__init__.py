"""Overlay of the reference's vendored `openfold` package (see framedipt_b200/dropin.py)."""
from framedipt_b200 import dropin as _d

__path__ = _d.overlay_path(__file__, "openfold")
