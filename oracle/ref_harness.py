"""TEST INFRASTRUCTURE ONLY — loader for the *unmodified* reference (instadeepai/FrameDiPT).

Imports the reference from ``/root/reference`` (or ``$FRAMEDIPT_REF``) behind stub modules
for import-time dependencies that are absent from this image (omegaconf, dm-tree, Bio,
ml_collections, GPUtil, hydra).  It exists so that ``oracle/make_golden.py`` can run the real
reference in the build container and commit its outputs as fixtures under ``tests/golden/``.
The reference tree does not exist on the GPU box; nothing on the product path may import this.
"""
from __future__ import annotations

import importlib.machinery
import os
import sys
import types


class AttrDict(dict):
    """Stand-in for omegaconf.DictConfig: attribute access over a nested dict."""

    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError as e:  # pragma: no cover
            raise AttributeError(k) from e
        return v

    def __setattr__(self, k, v):
        self[k] = v


def to_attr(d):
    if isinstance(d, dict):
        return AttrDict({k: to_attr(v) for k, v in d.items()})
    return d


def _stub(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, None)
    m.__path__ = []  # behave like a package
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def _map_structure(fn, *structs):
    s0 = structs[0]
    if isinstance(s0, dict):
        return {k: _map_structure(fn, *[s[k] for s in structs]) for k in s0}
    if isinstance(s0, (list, tuple)):
        return type(s0)(_map_structure(fn, *xs) for xs in zip(*structs))
    return fn(*structs)


class _FieldRef:
    """ml_collections.FieldReference stand-in (only arithmetic at import time is needed)."""

    def __init__(self, v, **_):
        self.v = v

    def get(self):
        return self.v

    def _b(self, o):
        return o.v if isinstance(o, _FieldRef) else o

    def __mul__(self, o): return _FieldRef(self.v * self._b(o))
    __rmul__ = __mul__
    def __add__(self, o): return _FieldRef(self.v + self._b(o))
    __radd__ = __add__
    def __floordiv__(self, o): return _FieldRef(self.v // self._b(o))
    def __truediv__(self, o): return _FieldRef(self.v / self._b(o))
    def __sub__(self, o): return _FieldRef(self.v - self._b(o))


def install_stubs() -> None:
    if "omegaconf" not in sys.modules:
        try:
            import omegaconf  # noqa: F401
        except ImportError:
            class _OC:
                @staticmethod
                def create(d):
                    return to_attr(d)

                @staticmethod
                def to_container(d, **_):
                    return dict(d)

                @staticmethod
                def set_struct(*_a, **_k):
                    return None

                @staticmethod
                def merge(*ds):
                    out = AttrDict()
                    for d in ds:
                        out.update(d)
                    return out

            _stub("omegaconf", DictConfig=AttrDict, OmegaConf=_OC, ListConfig=list)
    for name, attrs in [
        ("tree", dict(map_structure=_map_structure)),
        ("GPUtil", dict(getAvailable=lambda **_: [])),
        ("hydra", dict(main=lambda **_: (lambda f: f))),
        ("hydra.core", {}),
        ("hydra.core.hydra_config", dict(HydraConfig=type("HydraConfig", (), {"initialized": staticmethod(lambda: False)}))),
        ("ml_collections", dict(ConfigDict=AttrDict, FieldReference=_FieldRef)),
        ("Bio", {}),
        ("Bio.PDB", dict(MMCIFParser=object, Model=object, Structure=object, PDBParser=object, PDBIO=object, Chain=object)),
        ("Bio.PDB.Chain", dict(Chain=object)),
        ("Bio.PDB.Model", dict(Model=object)),
        ("Bio.PDB.Structure", dict(Structure=object)),
        ("Bio.PDB.MMCIFParser", dict(MMCIFParser=object)),
        ("Bio.PDB.PDBParser", dict(PDBParser=object)),
        ("Bio.PDB.PDBIO", dict(PDBIO=object)),
    ]:
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                _stub(name, **attrs)


def ref_root() -> str | None:
    for cand in (os.environ.get("FRAMEDIPT_REF"), "/root/reference"):
        if cand and os.path.isdir(os.path.join(cand, "framedipt")):
            return cand
    return None


def load_reference():
    """Returns (score_network, se3_diffuser, experiments.utils, rigid_utils, all_atom) modules."""
    root = ref_root()
    if root is None:
        raise RuntimeError("reference tree not available (expected /root/reference)")
    install_stubs()
    if root not in sys.path:
        sys.path.insert(0, root)
    from framedipt.model import score_network  # type: ignore
    from framedipt.diffusion import se3_diffuser  # type: ignore
    from experiments import utils as exp_utils  # type: ignore
    from openfold.utils import rigid_utils  # type: ignore
    from framedipt.protein import all_atom  # type: ignore
    return score_network, se3_diffuser, exp_utils, rigid_utils, all_atom


def default_conf(cache_dir: str = "/tmp/fdpt_igso3_cache", input_aatype: bool = True, seed: int = 123,
                 num_sigma: int = 1000, num_omega: int = 1000):
    """config/base.yaml:33-79 values as an attribute dict."""
    return to_attr({
        "diffuser": {
            "diffuse_trans": True, "diffuse_rot": True,
            "r3": {"min_b": 0.1, "max_b": 20.0, "coordinate_scaling": 0.1, "seed": seed},
            "so3": {"num_omega": num_omega, "num_sigma": num_sigma, "min_sigma": 0.1, "max_sigma": 1.5,
                    "schedule": "logarithmic", "cache_dir": cache_dir, "use_cached_score": False, "seed": seed},
        },
        "model": {
            "input_aatype": input_aatype, "node_embed_size": 256, "edge_embed_size": 128, "dropout": 0.0,
            "embed": {"index_embed_size": 32, "aatype_embed_size": 64, "embed_self_conditioning": True,
                      "num_bins": 22, "min_bin": 1e-5, "max_bin": 20.0},
            "ipa": {"c_s": 256, "c_z": 128, "c_hidden": 256, "c_skip": 64, "no_heads": 8, "no_qk_points": 8,
                    "no_v_points": 12, "seq_tfmr_num_heads": 4, "seq_tfmr_num_layers": 2, "num_blocks": 4,
                    "coordinate_scaling": 0.1},
        },
    })
