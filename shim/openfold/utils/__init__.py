"""Overlay of `openfold.utils` (see framedipt_b200/dropin.py)."""
from framedipt_b200 import dropin as _d

__path__ = _d.overlay_path(__file__, "openfold/utils")
