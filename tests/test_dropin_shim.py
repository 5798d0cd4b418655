"""Drop-in boundary (BASELINE.json north star: "drops into inference.py unchanged").

CPU part (build container, where the reference checkout is mounted): the reference's own, UNMODIFIED experiments/inference.py is imported
on top of the overlay `shim/` and its `_load_ckpt` / `create_dataset` / sampling loop run up to the `inference_fn` call, which the
B200 path refuses loudly without a GPU.  GPU part (no reference checkout there): the same call sequence through the overlay's module
paths, end to end."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ref_root():
    for c in (os.environ.get("FRAMEDIPT_REF"), "/root/reference"):
        if c and os.path.isfile(os.path.join(c, "experiments", "inference.py")):
            return c
    return None


def test_reference_inference_py_runs_unchanged_on_the_overlay():
    ref = _ref_root()
    if ref is None:
        pytest.skip("reference checkout not available on this box")
    env = dict(os.environ, PYTHONPATH="")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "dropin_driver.py"), ref, "cpu"], capture_output=True, text=True, env=env,
                       timeout=600)
    lines = [l for l in p.stdout.splitlines() if l.startswith("DROPIN ")]
    assert lines, p.stderr[-2000:]
    r = json.loads(lines[-1][7:])
    assert r["inference_py"] == os.path.join(ref, "experiments", "inference.py")  # the reference's file, not a copy
    assert r["score_network_is_ours"] and r["se3_diffuser_is_ours"] and r["inference_fn_is_ours"] and r["logp_is_ours"]
    assert r["model_class"] == "framedipt_b200.score_network.ScoreNetwork" and r["diffuser_class"] == "framedipt_b200.se3_diffuser.SE3Diffuser"
    assert r["params_loaded"] == 17_446_190  # de-novo variant: load_state_dict through Inference._load_ckpt (module. prefix stripped there)
    assert r["sampler_class"].endswith("UnconditionalSampler")
    # host-only helpers still come from the reference's own files
    assert r["other_helpers_module"] == "_framedipt_ref_experiments_utils" and r["rigid_module"] == "_framedipt_ref_rigid_utils"
    # the sampling loop reached OUR inference_fn from the reference's run_unconditional_sampling; without a GPU it refuses (no CPU fallback)
    assert r["raised_in"][-2:] == ["inference.py:run_unconditional_sampling", "inference.py:inference_fn"]
    assert "FdptError" in r["sampling"] and "CUDA" in r["sampling"]


def test_overlay_imports_without_the_reference_checkout():
    """On a box without the reference (the GPU box) every module path of the overlay still imports and serves the sampler surface."""
    code = ("import sys; sys.path[:0]=[%r, %r]; import os; os.environ.pop('FRAMEDIPT_REF', None);"
            "from framedipt.model import score_network; from framedipt.diffusion import se3_diffuser; from experiments import utils, sampler;"
            "from openfold.utils import rigid_utils; from framedipt.data import utils as du; import framedipt;"
            "print(score_network.ScoreNetwork.__module__, se3_diffuser.SE3Diffuser.__module__, utils.inference_fn.__module__,"
            " sampler.UnconditionalSampler.__module__, rigid_utils.Rigid.__module__, du.pad_feats.__module__, framedipt.RESIDUE_GAP)") % (
        os.path.join(ROOT, "shim"), ROOT)
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=dict(os.environ, PYTHONPATH=""), cwd="/tmp", timeout=300)
    assert p.returncode == 0, p.stderr[-1500:]
    assert p.stdout.split() == ["framedipt_b200.score_network", "framedipt_b200.se3_diffuser", "framedipt_b200.inference", "framedipt_b200.sampler",
                                "framedipt_b200.rigid", "framedipt_b200.sampler", "200"]


@pytest.mark.gpu
def test_inference_py_call_sequence_through_the_overlay(tmp_path):
    """The call sequence of experiments/inference.py (107-161 `_load_ckpt`, 163-183 `create_dataset`, 195-240 the sampling loop) written
    against the REFERENCE's module paths, resolved by the overlay: checkpoint -> model -> sampler -> inference_fn -> PDB of the sample."""
    code = r'''
import sys, json
sys.path[:0] = [%(shim)r, %(root)r]
import numpy as np, torch
from experiments import sampler, utils as exp_utils
from framedipt.data import utils as data_utils
from framedipt.diffusion import se3_diffuser
from framedipt.model import score_network
from openfold.utils import rigid_utils
from framedipt_b200.config import default_conf, to_attr
from framedipt_b200.params import synthetic_state_dict
from framedipt_b200.pdb import write_prot_to_pdb

conf = default_conf(input_aatype=False)
ckpt = %(ckpt)r
data_utils.write_pkl(ckpt, {"conf": dict(conf), "model": {"module." + k: v for k, v in synthetic_state_dict(0, with_aatype=False).items()}}, use_torch=True)
device = "cuda:0"
weights_pkl = data_utils.read_pkl(ckpt, use_torch=True, map_location=device)
conf.diffuser.so3.seed = conf.diffuser.r3.seed = 123
diffuser = se3_diffuser.SE3Diffuser(conf.diffuser)
model = score_network.ScoreNetwork(conf.model, diffuser, inpainting=False)
model.load_state_dict({k.replace("module.", ""): v for k, v in weights_pkl["model"].items()})
model = model.to(device)
model.eval()
ds = sampler.UnconditionalSampler(cfg=to_attr({"min_length": 40, "max_length": 40, "length_step": 1, "samples_per_length": 2}), diffuser=diffuser, device=device)
out = []
for sample_length, sample_i, sample_feats in ds:
    so = exp_utils.inference_fn(model=model, diffuser=diffuser, data_init=sample_feats, num_t=4, min_t=0.01, noise_scale=0.1, aux_traj=True,
                                embed_self_conditioning=True, inpainting=False, input_aatype=False)
    so = {k: v[:, 0] for k, v in so.items()}
    rig = rigid_utils.Rigid.from_tensor_7(torch.tensor(so["rigid_traj"][0]))
    path = write_prot_to_pdb(so["prot_traj"][0], %(out)r + f"/sample_{sample_i}.pdb", no_indexing=True)
    out.append({"len": int(sample_length), "i": int(sample_i), "prot": list(so["prot_traj"].shape), "rigid": list(so["rigid_traj"].shape),
                "finite": bool(np.isfinite(so["prot_traj"]).all()), "trans_ok": bool(np.allclose(rig.get_trans().numpy(), so["prot_traj"][0][:, 1], atol=1e-4)),
                "pdb": str(path)})
print("SEQ " + json.dumps(out))
''' % {"shim": os.path.join(ROOT, "shim"), "root": ROOT, "ckpt": str(tmp_path / "denovo.pth"), "out": str(tmp_path)}
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=dict(os.environ, PYTHONPATH=""), cwd=str(tmp_path), timeout=900)
    lines = [l for l in p.stdout.splitlines() if l.startswith("SEQ ")]
    assert lines, (p.stdout[-500:], p.stderr[-2500:])
    r = json.loads(lines[-1][4:])
    assert [x["i"] for x in r] == [0, 1]
    for x in r:
        assert x["len"] == 40 and x["prot"] == [4, 40, 37, 3] and x["rigid"] == [5, 40, 7] and x["finite"] and x["trans_ok"]
        assert os.path.getsize(x["pdb"]) > 40 * 4 * 80
