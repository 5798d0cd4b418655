"""Overlay of the reference's `framedipt` package (see framedipt_b200/dropin.py)."""
from framedipt_b200 import dropin as _d

__path__ = _d.overlay_path(__file__, "framedipt")
RESIDUE_GAP = 200  # framedipt/__init__.py:3
