"""CPU tests: host logic, C-ABI exports, synthetic inputs, schedule."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from util import rot_angle_between

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "fdpt.h")).read()
    names = set(re.findall(r"\b(fdpt_[a-z_0-9]+)\s*\(", hdr))
    assert len(names) >= 15
    lib_path = os.path.join(ROOT, "framedipt_b200", "libfdpt.so")
    if not os.path.exists(lib_path):
        import __graft_entry__ as ge
        ge.build()
    lib = ctypes.CDLL(lib_path)
    for n in names:
        assert hasattr(lib, n), n
    lib.fdpt_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.fdpt_version()


def test_param_specs_match_reference_inventory():
    from framedipt_b200.params import param_specs, synthetic_state_dict

    sd = synthetic_state_dict(0)
    assert sum(v.numel() for v in sd.values()) == 17_456_942
    assert len(sd) == len(param_specs())
    sd2 = synthetic_state_dict(0)
    assert all(torch.equal(sd[k], sd2[k]) for k in sd)


def test_sample_ref_matches_reference(golden_dir):
    from framedipt_b200 import SE3Diffuser, synthetic
    from framedipt_b200.config import default_conf

    g = np.load(os.path.join(golden_dir, "sample_ref.npz"))
    diffuser = SE3Diffuser(default_conf().diffuser)
    np.random.seed(123)
    f = synthetic.make_features(synthetic.WORKLOADS["cfg1_monomer64"], diffuser, seed=0)
    r, ref = f["rigids_t"].numpy(), g["inpaint_rigids_t"]
    assert np.abs(r[..., 4:] - ref[..., 4:]).max() < 1e-5
    assert rot_angle_between(r[..., :4], ref[..., :4]).max() < 1e-5
    f2 = synthetic.make_features(synthetic.Workload("denovo32", 2, (32,), (), 10, de_novo=True), diffuser, seed=0)
    r, ref = f2["rigids_t"].numpy(), g["denovo_rigids_t"]
    assert np.abs(r[..., 4:] - ref[..., 4:]).max() < 1e-5
    assert rot_angle_between(r[..., :4], ref[..., :4]).max() < 1e-5


def test_schedule_matches_oracle():
    from framedipt_b200 import SE3Diffuser
    from framedipt_b200.config import default_conf
    from oracle import framedipt_oracle as orc

    d = SE3Diffuser(default_conf().diffuser)
    for num_t in (50, 100, 500):
        for t in np.linspace(0.01, 1.0, num_t):
            row = d.step_scalars(float(t), 1 / num_t, 0.1)
            assert row[1] == orc.sigma_grid_value(np.array(np.float32(t)))
            g = orc.so3_diffusion_coef(t)
            assert abs(row[2] - g * g / num_t) < 1e-15 and abs(row[4] - (0.1 + t * 19.9)) < 1e-15
    with pytest.raises(ValueError):
        d._so3_diffuser.sigma(1.5)


def test_sample_ref_errors():
    from framedipt_b200 import SE3Diffuser
    from framedipt_b200.config import default_conf
    from framedipt_b200.rigid import Rigid

    d = SE3Diffuser(default_conf().diffuser)
    with pytest.raises(ValueError):
        d.sample_ref(n_samples=4, diffuse_mask=np.ones(4))
    with pytest.raises(ValueError):
        d.sample_ref(n_samples=4, impute=Rigid.identity((5,)), diffuse_mask=np.ones(4))


def test_host_embeddings_match_oracle():
    from framedipt_b200 import runtime
    from oracle import framedipt_oracle as orc

    idx = torch.tensor([[0, 1, 5, 230, 1023], [3, 7, 300, 500, 900]])
    assert torch.equal(runtime.index_embedding(idx), orc.index_embedding(idx).float())
    t = torch.tensor([0.01, 0.37, 1.0])
    assert torch.equal(runtime.timestep_embedding(t), orc.timestep_embedding(t))


def test_workloads_cover_baseline_configs():
    from framedipt_b200 import synthetic

    w = synthetic.WORKLOADS
    assert w["cfg1_monomer64"].n_res == 64 and w["cfg2_tcr350"].n_res == 350 and w["cfg3_denovo256"].batch == 64
    assert w["cfg4_tcrpmhc800"].n_res == 800 and w["cfg5_sweep1024"].n_res == 1024
    st = synthetic.static_features(w["cfg2_tcr350"], 0)
    assert st["seq_idx"][170] == 170 + 200  # RESIDUE_GAP between chains
    assert st["fixed_mask"].sum() == 350 - 24


def test_confidence_score_transition_densities_match_reference(golden_dir):
    """SE3Diffuser.forward / log_prob_forward / log_prob_backward (the EigenFold confidence-score pieces, se3_diffuser.py:50-196)
    against values produced by the unmodified reference (oracle/make_golden_logp.py): with the same legacy numpy seed the noised
    frames are identical, and both log-densities agree to the last bit (same arithmetic types, same summation order)."""
    import copy

    from framedipt_b200 import SE3Diffuser
    from framedipt_b200.config import default_conf
    from framedipt_b200.rigid import rotmats_to_rigid

    g = np.load(os.path.join(golden_dir, "logp_small.npz"))
    d = SE3Diffuser(default_conf().diffuser)
    mask = ((1 - g["in_fixed_mask"]) * g["in_res_mask"])[0].astype(np.float64)
    assert 0 < mask.sum() < mask.size
    for ci in range(int(g["n_cases"])):
        t_1, dt, t, lpf, lpb = (float(v) for v in g[f"c{ci}_scalars"])
        a = rotmats_to_rigid(g[f"c{ci}_a_rot"], g[f"c{ci}_a_trans"])
        b = rotmats_to_rigid(g[f"c{ci}_b_rot"], g[f"c{ci}_b_trans"])
        np.random.seed(1000 + ci)
        b2 = d.forward(copy.deepcopy(a), t_1, dt, mask)
        assert np.array_equal(b2.get_rots().get_rot_mats().numpy(), g[f"c{ci}_b_rot"])
        assert np.array_equal(b2.get_trans().numpy(), g[f"c{ci}_b_trans"])
        fixed = mask == 0  # fixed residues are not moved
        assert np.array_equal(g[f"c{ci}_b_trans"][fixed], g[f"c{ci}_a_trans"][fixed])
        assert d.log_prob_forward(b, a, t_1, dt, mask) == lpf
        assert d.log_prob_backward(b, a, g[f"c{ci}_trans_score"], g[f"c{ci}_rot_score"], t, dt, mask) == lpb
    # the densities are ordinary Gaussians: moving x(t) away from the forward mean lowers log q
    a = rotmats_to_rigid(g["c1_a_rot"], g["c1_a_trans"])
    far = rotmats_to_rigid(g["c1_b_rot"], g["c1_b_trans"] + 5.0 * mask[:, None].astype(np.float32))
    assert d.log_prob_forward(far, a, 0.3, 0.1, mask) < float(g["c1_scalars"][3])


def test_backbone_constants_match_reference_tables(golden_dir):
    """framedipt_b200/backbone_constants.py against the tables derived from the reference's residue_constants
    (oracle/make_constants.py -> tests/golden/backbone_tables.npz), all 20 residue types."""
    from framedipt_b200 import backbone_constants as bc

    g = np.load(os.path.join(golden_dir, "backbone_tables.npz"))
    assert bc.IDEAL_BB_POS.shape == (20, 5, 3) and np.array_equal(np.asarray(bc.IDEAL_BB_POS, np.float32), g["ideal"])
    assert bc.PSI_DEFAULT_FRAME.shape == (20, 4, 4) and np.array_equal(np.asarray(bc.PSI_DEFAULT_FRAME, np.float32), g["psi_frame"])
    assert np.array_equal(np.asarray(bc.BB_ATOM_MASK, np.float32), g["mask"])
    assert g["mask"][7, 4] == 0 and g["mask"].sum() == 99  # GLY is the only type without CB


def test_score_norm_rows_match_reference_table(golden_dir):
    """Host builder of the cached IGSO(3) score-norm table (so3.use_cached_score=True) vs rows of the reference's `_score_norms`."""
    from framedipt_b200 import SE3Diffuser
    from framedipt_b200.config import default_conf

    g = np.load(os.path.join(golden_dir, "config_variants.npz"))
    so3 = SE3Diffuser(default_conf().diffuser)._so3_diffuser
    for r, row in zip(g["cached_table_rows"], g["cached_table"]):
        assert np.abs(so3.score_norm_row(int(r)) - row).max() <= 1e-9 * np.abs(row).max()


def test_sigma_index_is_the_same_for_float32_and_float64_t():
    """the reference pins numpy 1.22 (value-based casting: sigma(t) of a float32 t stays float32), the fixtures were
    produced under numpy 2 (float64).  On the schedules actually used the grid index does not depend on that."""
    from framedipt_b200 import SE3Diffuser
    from framedipt_b200.config import default_conf

    so3 = SE3Diffuser(default_conf().diffuser)._so3_diffuser
    for num_t in (50, 100, 200, 500):
        t = np.linspace(0.01, 1.0, num_t)
        i64 = so3.t_to_idx(t)
        i32 = so3.t_to_idx(t.astype(np.float32).astype(np.float64))
        sig32 = np.log(t.astype(np.float32) * np.float32(np.exp(1.5)) + (np.float32(1) - t.astype(np.float32)) * np.float32(np.exp(0.1)))
        i32b = np.digitize(sig32.astype(np.float64), so3.discrete_sigma) - 1  # all-float32 evaluation (numpy 1.22 semantics)
        assert np.array_equal(i64, i32), num_t
        assert (i64 != i32b).sum() <= 1, (num_t, np.nonzero(i64 != i32b))  # only an exact bin edge (t = 1.0) may flip


def test_samplers_keep_the_reference_item_contract(golden_dir):
    """(key, sample_i, feats) items with batch dim 1 (experiments/sampler.py:121-135, 267-354); the de-novo sampler consumes the
    legacy numpy stream exactly like the reference (fixture written by the unmodified sample_ref)."""
    from framedipt_b200 import SE3Diffuser, synthetic
    from framedipt_b200.config import default_conf, to_attr
    from framedipt_b200.sampler import SyntheticConditionalSampler, UnconditionalSampler

    g = np.load(os.path.join(golden_dir, "sample_ref.npz"))
    diffuser = SE3Diffuser(default_conf().diffuser)
    # the fixture's stream: seed, one inpainting sample (N=64), then two de-novo samples (N=32)
    cs = SyntheticConditionalSampler(to_attr({"workloads": ["cfg1_monomer64"], "samples": 2, "seed": 0}), diffuser, "cpu")
    assert len(cs) == 2
    np.random.seed(123)
    name, si, f = cs[0]
    assert (name, si) == ("cfg1_monomer64", 0)
    for key, shp in {"aatype": (1, 64), "seq_idx": (1, 64), "chain_idx": (1, 64), "res_mask": (1, 64), "fixed_mask": (1, 64),
                     "torsion_angles_sin_cos": (1, 64, 7, 2), "rigids_0": (1, 64, 7), "sc_ca_t": (1, 64, 3), "rigids_t": (1, 64, 7), "t": (1,)}.items():
        assert tuple(f[key].shape) == shp, key
    assert np.abs(f["rigids_t"][0, :, 4:].numpy() - g["inpaint_rigids_t"][0, :, 4:]).max() < 1e-5
    ds = UnconditionalSampler(to_attr({"min_length": 32, "max_length": 40, "length_step": 8, "samples_per_length": 2}), diffuser, "cpu")
    assert len(ds) == 4 and list(ds.all_sampling_lengths) == [32, 32, 40, 40]
    items = [ds[0], ds[1]]
    for k, (n, i, f) in enumerate(items):
        assert (n, i) == (32, k)
        assert f["rigids_t"].shape == (1, 32, 7) and f["res_mask"].shape == (1, 32) and f["torsion_angles_sin_cos"].shape == (1, 32, 7, 2)
        assert torch.equal(f["seq_idx"][0], torch.arange(1, 33))
        assert np.abs(f["rigids_t"][0, :, 4:].numpy() - g["denovo_rigids_t"][k, :, 4:]).max() < 1e-5
    assert len(list(iter(cs))) == 2  # iterable Dataset protocol used by `for ... in self.sampler`


def test_pad_feats_matches_reference(golden_dir):
    """pad_feats / pad_rigid (framedipt/data/utils.py:311-339): the padded mixed-length batch of the fixture was built by the
    reference's own functions from the same two structures."""
    from framedipt_b200 import SE3Diffuser, synthetic
    from framedipt_b200.config import default_conf
    from framedipt_b200.sampler import pad_feats

    g = np.load(os.path.join(golden_dir, "config_variants.npz"))
    diffuser = SE3Diffuser(default_conf().diffuser)
    np.random.seed(123)
    fa = synthetic.make_features(synthetic.Workload("mixA", 1, (12, 8), ((3, 8),), 6), diffuser, seed=41)
    fb = synthetic.make_features(synthetic.Workload("mixB", 1, (31,), ((10, 19),), 6), diffuser, seed=42)
    pa = pad_feats({k: v[0] for k, v in fa.items()}, 31, use_torch=True)
    pb = pad_feats({k: v[0] for k, v in fb.items()}, 31, use_torch=True)
    for k in pa:
        if k == "t":
            continue
        ours = torch.stack([pa[k], pb[k]]).numpy()
        ref = g[f"mixed_in_{k}"]
        assert ours.shape == ref.shape, k
        if k == "rigids_t":
            assert np.abs(ours[..., 4:] - ref[..., 4:]).max() < 1e-5 and rot_angle_between(ours[..., :4], ref[..., :4]).max() < 1e-5
        else:
            assert np.array_equal(ours, ref), k
    with pytest.raises(ValueError):
        pad_feats({"res_mask": torch.ones(40)}, 31, use_torch=True)
