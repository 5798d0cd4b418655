"""TEST INFRASTRUCTURE — runs the reference's OWN `experiments/inference.py` (unmodified, imported from the reference checkout) on top of
the drop-in overlay `shim/`: `Inference._load_ckpt` (inference.py:107-161), `create_dataset` (163-183) and the sampling loop up to and
including its `exp_utils.inference_fn(...)` call (211-222).  Third-party packages the reference imports at module level but that are
absent from this image (esm, biotite, pdbfixer, openmm, mdtraj, tmtools, anarci, hydra, omegaconf, Bio ...) are replaced by inert
stand-ins; none of them is on the code path exercised here.  Prints one JSON line.

    python tests/dropin_driver.py /path/to/FrameDiPT [cuda|cpu]
"""
import importlib.abc
import importlib.machinery
import json
import os
import sys
import tempfile
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ref = sys.argv[1]
device = sys.argv[2] if len(sys.argv) > 2 else "cpu"
sys.path[:0] = [os.path.join(ROOT, "shim"), ROOT, ref]  # the overlay goes AHEAD of the reference checkout

import numpy as np  # noqa: E402
import pandas  # noqa: E402,F401
import scipy  # noqa: E402,F401
import torch  # noqa: E402

from oracle import ref_harness as rh  # noqa: E402

rh.install_stubs()
_torch_load = torch.load
torch.load = lambda *a, **k: _torch_load(*a, **{**k, "weights_only": False})  # the reference pins torch 1.13 (pickled config object inside)


class _Dummy:
    """absorbs any use made of an absent dependency at import time: attribute access, calls, arithmetic, subclassing"""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Dummy()

    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        return _Dummy()

    def _op(self, *a):
        return _Dummy()

    __add__ = __radd__ = __sub__ = __rsub__ = __mul__ = __rmul__ = __truediv__ = __rtruediv__ = __pow__ = __rpow__ = __neg__ = _op

    def __mro_entries__(self, bases):
        return (object,)

    def __iter__(self):
        return iter(())


class _Anything(types.ModuleType):
    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        v = _Dummy()
        setattr(self, k, v)
        return v


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    ABSENT = ("pdbfixer", "openmm", "simtk", "esm", "biotite", "mdtraj", "tmtools", "anarci", "matplotlib", "wandb", "Bio", "GPUtil", "hydra",
              "ml_collections", "deepspeed", "dllogger", "seaborn", "plotly", "py3Dmol", "pytorch_lightning")

    def find_spec(self, name, path=None, target=None):
        if name.split(".")[0] not in self.ABSENT:
            return None
        return importlib.machinery.ModuleSpec(name, self, is_package=True)

    def create_module(self, spec):
        m = _Anything(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


sys.meta_path.append(_StubFinder())

import experiments.inference as inf  # noqa: E402  (the reference's file: the overlay has no experiments/inference.py)

import framedipt_b200.inference as our_inf  # noqa: E402
import framedipt_b200.score_network as our_sn  # noqa: E402
import framedipt_b200.se3_diffuser as our_se3  # noqa: E402
from framedipt_b200.params import synthetic_state_dict  # noqa: E402

res = {"inference_py": inf.__file__, "score_network_is_ours": inf.score_network.ScoreNetwork is our_sn.ScoreNetwork,
       "se3_diffuser_is_ours": inf.se3_diffuser.SE3Diffuser is our_se3.SE3Diffuser,
       "inference_fn_is_ours": inf.exp_utils.inference_fn is our_inf.inference_fn,
       "logp_is_ours": inf.logp_confidence_score is our_inf.logp_confidence_score,
       "sampler_module": inf.sampler.UnconditionalSampler.__module__, "rigid_module": inf.rigid_utils.Rigid.__module__,
       "other_helpers_module": inf.exp_utils.get_atom_positions_from_rigids.__module__}

tmp = tempfile.mkdtemp()
conf = rh.default_conf(cache_dir=os.path.join(tmp, "igso3"), input_aatype=False, seed=123)
ckpt = os.path.join(tmp, "denovo.pth")
sd = synthetic_state_dict(0, with_aatype=False)
torch.save({"conf": conf, "model": {"module." + k: v for k, v in sd.items()}}, ckpt)

cfg = rh.to_attr({
    "model": dict(conf.model), "diffuser": dict(conf.diffuser),
    "inference": {"seed": 123, "inpainting": False, "input_aatype": False, "diffusion": {"num_t": 3, "min_t": 0.01, "noise_scale": 0.1},
                  "samples": {"min_length": 32, "max_length": 32, "length_step": 1, "samples_per_length": 1}},
})
obj = object.__new__(inf.Inference)  # __init__ needs Hydra run directories, ESMFold and a ProteinMPNN checkout: not on this path
obj._cfg = cfg
obj.device = device
obj._load_ckpt(ckpt, None)  # inference.py:107-161, unchanged
res["model_class"] = type(obj.model).__module__ + "." + type(obj.model).__name__
res["diffuser_class"] = type(obj.diffuser).__module__ + "." + type(obj.diffuser).__name__
res["params_loaded"] = int(sum(p.numel() for p in obj.model.parameters()))
res["model_device"] = str(next(obj.model.parameters()).device)
obj.create_dataset()  # inference.py:163-183
res["sampler_class"] = type(obj.sampler).__module__ + "." + type(obj.sampler).__name__
import pathlib  # noqa: E402

obj.output_dir = pathlib.Path(tmp) / "out"
np.random.seed(123)
try:
    obj.run_unconditional_sampling()  # inference.py:195-240: iterates the sampler, calls exp_utils.inference_fn(...)
    res["sampling"] = "completed"
except Exception as e:  # on a box without a GPU the B200 path refuses loudly at exactly that call; with one, the steps after it need ESMFold
    res["sampling"] = f"{type(e).__module__}.{type(e).__name__}: {str(e)[:160]}"
    import traceback

    tb = traceback.extract_tb(e.__traceback__)
    res["raised_in"] = [f"{os.path.basename(f.filename)}:{f.name}" for f in tb][-4:]
pdbs = sorted(str(p.relative_to(obj.output_dir)) for p in obj.output_dir.rglob("*.pdb")) if obj.output_dir.exists() else []
res["pdb_files"] = pdbs
print("DROPIN " + json.dumps(res))
