"""ctypes binding of libfdpt.so (include/fdpt.h) + host-side feature preparation.

PyTorch is used here for device memory and streams only; all math of the hot path runs in the
hand-written sm_100a kernels of ``libfdpt.so``.  There is no CPU fallback: importing this module
without the built library, or calling it without a CUDA device, raises.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Any

import numpy as np
import torch

from .params import ModelDims, param_specs

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfdpt.so")
SCHED_COLS = 10  # FDPT_SCHED_COLS (include/fdpt.h)


class FdptError(RuntimeError):
    pass


class _Config(C.Structure):
    _fields_ = [
        ("c_s", C.c_int32), ("c_z", C.c_int32), ("c_hidden", C.c_int32), ("c_skip", C.c_int32),
        ("no_heads", C.c_int32), ("no_qk_points", C.c_int32), ("no_v_points", C.c_int32), ("num_blocks", C.c_int32),
        ("index_embed_size", C.c_int32), ("num_bins", C.c_int32), ("min_bin", C.c_float), ("max_bin", C.c_float),
        ("seq_tfmr_num_heads", C.c_int32), ("seq_tfmr_num_layers", C.c_int32), ("coordinate_scaling", C.c_float),
        ("with_aatype", C.c_int32), ("r3_min_b", C.c_double), ("r3_max_b", C.c_double),
        ("r3_coordinate_scaling", C.c_float), ("embed_self_conditioning", C.c_int32),
    ]


class _Feats(C.Structure):
    _fields_ = [
        ("rigids_t", C.c_void_p), ("sc_ca_t", C.c_void_p), ("res_mask", C.c_void_p), ("fixed_mask", C.c_void_p),
        ("seq_idx", C.c_void_p), ("aatype", C.c_void_p), ("gt_psi", C.c_void_p), ("idx_emb", C.c_void_p),
        ("rel_emb", C.c_void_p), ("rel_min", C.c_int32), ("rel_count", C.c_int32), ("t_emb", C.c_void_p),
        ("t_emb_eps", C.c_void_p), ("t32", C.c_void_p), ("sigma", C.c_void_p), ("sigma_idx", C.c_void_p),
        ("aatype_bb", C.c_void_p), ("aatype_bb_given", C.c_int32),
    ]


class _Out(C.Structure):
    _fields_ = [("rigids", C.c_void_p), ("rot_score", C.c_void_p), ("trans_score", C.c_void_p), ("psi", C.c_void_p),
                ("atom37_bb", C.c_void_p)]


class _Traj(C.Structure):
    _fields_ = [("prot_traj", C.c_void_p), ("rigid_traj", C.c_void_p), ("trans_traj", C.c_void_p),
                ("rigid_0_traj", C.c_void_p), ("psi_pred", C.c_void_p), ("final_only", C.c_int32)]


_lib = None


def lib() -> C.CDLL:
    """Loads libfdpt.so; raises (no fallback) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FdptError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'`")
        L = C.CDLL(LIB_PATH)
        L.fdpt_last_error.restype = C.c_char_p
        L.fdpt_last_error.argtypes = [C.c_void_p]
        L.fdpt_version.restype = C.c_char_p
        L.fdpt_create.argtypes = [C.POINTER(_Config), C.c_int, C.POINTER(C.c_void_p)]
        L.fdpt_destroy.argtypes = [C.c_void_p]
        L.fdpt_load_param.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.POINTER(C.c_int64), C.c_int]
        L.fdpt_finalize_params.argtypes = [C.c_void_p]
        L.fdpt_num_params_expected.argtypes = [C.c_void_p]
        L.fdpt_reserve.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.fdpt_workspace_bytes.restype = C.c_int64
        L.fdpt_workspace_bytes.argtypes = [C.c_void_p]
        L.fdpt_launch_count.restype = C.c_int64
        L.fdpt_launch_count.argtypes = [C.c_void_p]
        L.fdpt_stat.restype = C.c_int64
        L.fdpt_stat.argtypes = [C.c_void_p, C.c_int]
        L.fdpt_forward.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(_Feats), C.POINTER(_Out), C.c_void_p]
        L.fdpt_reverse.argtypes = [C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 6 + [C.POINTER(C.c_double), C.c_int, C.c_int, C.c_int,
                                                                                    C.c_void_p, C.c_void_p]
        L.fdpt_backbone.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.fdpt_rot_score.argtypes = [C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 6
        L.fdpt_trans_score.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                       C.c_void_p, C.c_void_p]
        L.fdpt_sample.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(_Feats), C.c_int, C.POINTER(C.c_double), C.c_void_p,
                                  C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(_Traj), C.c_void_p]
        L.fdpt_set_progress_chunk.argtypes = [C.c_void_p, C.c_int]
        L.fdpt_wait_step.argtypes = [C.c_void_p, C.c_int]
        L.fdpt_set_score_table.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.fdpt_rot_score_idx.argtypes = [C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 6
        L.fdpt_sample_ref.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                      C.c_uint64, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.fdpt_tmem_a_selftest.argtypes = [C.c_void_p] * 5
        L.fdpt_seq_tfmr.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 6
        L.fdpt_linear.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                  C.c_void_p]
        L.fdpt_tc_linear.argtypes = L.fdpt_linear.argtypes
        L.fdpt_matmul.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_longlong, C.c_void_p, C.c_int,
                                  C.c_longlong, C.c_int, C.c_float, C.c_void_p, C.c_int, C.c_longlong, C.c_void_p]
        L.fdpt_set_option.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.fdpt_to_pdb.restype = C.c_int64
        L.fdpt_to_pdb.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_int64]
        L.fdpt_bench_linear.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                        C.POINTER(C.c_float)]
        L.fdpt_debug_read.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.c_int]
        L.fdpt_ipa.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 7
        L.fdpt_edge_transition.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 5
        L.fdpt_profile_enable.argtypes = [C.c_void_p, C.c_int]
        L.fdpt_profile_read.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_double)]
        L.fdpt_embed.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(_Feats), C.c_void_p, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


# --------------------------------------------------------------------------------------------------------
# host-evaluated embeddings (fp32 trig arguments must be reproduced, so these tables are
# computed once on the host with the same torch-CPU expressions as the reference and the kernels only gather)
# --------------------------------------------------------------------------------------------------------
def index_embedding(indices: torch.Tensor, embed_size: int = 32, max_len: int = 2056) -> torch.Tensor:
    """get_index_embedding, framedipt/model/score_network.py:17-38 (int64 indices -> float32 [.., embed_size])."""
    indices = indices.to("cpu", torch.int64)
    k = torch.arange(embed_size // 2)
    arg = indices[..., None] * math.pi / (max_len ** (2 * k[None] / embed_size))
    return torch.cat([torch.sin(arg), torch.cos(arg)], -1).float()


def timestep_embedding(t: torch.Tensor, dim: int = 32, max_positions: int = 10000) -> torch.Tensor:
    """get_timestep_embedding, score_network.py:41-64 (t float32 [B])."""
    if t.dim() != 1:
        raise ValueError(f"timesteps should have 1D shape, got {t.shape}.")
    t = t.to("cpu", torch.float32) * max_positions
    half = dim // 2
    emb = torch.exp(torch.arange(half, dtype=torch.float32) * -(math.log(max_positions) / (half - 1)))
    emb = t.float()[:, None] * emb[None, :]
    return torch.cat([torch.sin(emb), torch.cos(emb)], 1)


def _ptr(t: torch.Tensor | None) -> int | None:
    return None if t is None else t.data_ptr()


def _dev(x: Any, device, dtype) -> torch.Tensor:
    return torch.as_tensor(x).to(device=device, dtype=dtype).contiguous()


class PreparedFeats:
    """Device-resident, dtype-normalised view of a reference feature dict."""

    def __init__(self, feats: dict, device, dims: ModelDims, with_aatype: bool, aatype_pre: torch.Tensor | None,
                 aatype_bb: torch.Tensor | None | bool = False):
        f32, i32 = torch.float32, torch.int32
        if with_aatype and aatype_pre is None:
            raise ValueError("When inpainting is True, aatype should be given, got None.")
        self.B, self.N = feats["res_mask"].shape
        self.rigids_t = _dev(feats["rigids_t"], device, f32)
        self.sc_ca_t = _dev(feats["sc_ca_t"], device, f32)
        self.res_mask = _dev(feats["res_mask"], device, f32)
        self.fixed_mask = _dev(feats["fixed_mask"], device, f32)
        seq = torch.as_tensor(feats["seq_idx"]).to("cpu", torch.int64)
        self.seq_idx = seq.to(device=device, dtype=i32).contiguous()
        self.aatype = _dev(aatype_pre, device, i32) if with_aatype else None
        # residue types of the trajectory's backbone atoms (inference_fn's own flags, experiments/utils.py:549-555); False = not given
        self.aatype_bb_given = aatype_bb is not False
        self.aatype_bb = _dev(aatype_bb, device, i32) if (self.aatype_bb_given and aatype_bb is not None) else None
        self.gt_psi = _dev(torch.as_tensor(feats["torsion_angles_sin_cos"])[..., 2, :], device, f32)
        self.idx_emb = index_embedding(seq, dims.index_embed_size).to(device).contiguous()
        lo = int((seq.min(-1).values - seq.max(-1).values).min())
        hi = int((seq.max(-1).values - seq.min(-1).values).max())
        self.rel_min, self.rel_count = lo, hi - lo + 1
        self.rel_emb = index_embedding(torch.arange(lo, hi + 1), dims.index_embed_size).to(device).contiguous()
        self.t_emb_eps = timestep_embedding(torch.tensor([1e-5]), dims.index_embed_size)[0].to(device).contiguous()
        self.t_emb = self.t32 = self.sigma = self.sigma_idx = None

    def struct(self) -> _Feats:
        return _Feats(_ptr(self.rigids_t), _ptr(self.sc_ca_t), _ptr(self.res_mask), _ptr(self.fixed_mask), _ptr(self.seq_idx),
                      _ptr(self.aatype), _ptr(self.gt_psi), _ptr(self.idx_emb), _ptr(self.rel_emb), self.rel_min, self.rel_count,
                      _ptr(self.t_emb), _ptr(self.t_emb_eps), _ptr(self.t32), _ptr(self.sigma), _ptr(self.sigma_idx),
                      _ptr(self.aatype_bb), int(self.aatype_bb_given))


class Context:
    """One libfdpt context (one GPU)."""

    def __init__(self, dims: ModelDims = ModelDims(), with_aatype: bool = True, device: int | torch.device = 0,
                 r3_min_b: float = 0.1, r3_max_b: float = 20.0, r3_coordinate_scaling: float = 0.1):
        if not torch.cuda.is_available():
            raise FdptError("framedipt_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device = torch.device("cuda", device if isinstance(device, int) else (device.index or 0))
        self.dims, self.with_aatype = dims, with_aatype
        cfg = _Config(dims.c_s, dims.c_z, dims.c_hidden, dims.c_skip, dims.no_heads, dims.no_qk_points, dims.no_v_points,
                      dims.num_blocks, dims.index_embed_size, dims.num_bins, dims.min_bin, dims.max_bin, dims.seq_tfmr_num_heads,
                      dims.seq_tfmr_num_layers, dims.coordinate_scaling, int(with_aatype), r3_min_b, r3_max_b,
                      r3_coordinate_scaling, int(dims.embed_self_conditioning))
        self.r3_key = (float(r3_min_b), float(r3_max_b), float(r3_coordinate_scaling))
        self._h = C.c_void_p()
        rc = lib().fdpt_create(C.byref(cfg), self.device.index, C.byref(self._h))
        if rc != 0:
            raise FdptError(f"fdpt_create failed ({rc}): unsupported configuration or CUDA error")
        self._params_loaded = False
        # A/B switches for the tools under tools/ (integrators never set these): FDPT_OPT_<n>=<value> -> fdpt_set_option(n, value)
        for k, v in os.environ.items():
            if k.startswith("FDPT_OPT_") and k[9:].isdigit():
                self.set_option(int(k[9:]), int(v))

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                lib().fdpt_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def _ck(self, rc: int):
        if rc != 0:
            msg = lib().fdpt_last_error(self._h).decode()
            raise FdptError(f"libfdpt error {rc}: {msg}")

    @property
    def stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def launch_count(self) -> int:
        return int(lib().fdpt_launch_count(self._h))

    def stat(self, which: int) -> int:
        """0: per-timestep graphs captured so far; 1: host microseconds spent inside the last fdpt_sample call."""
        return int(lib().fdpt_stat(self._h, which))

    def workspace_bytes(self) -> int:
        return int(lib().fdpt_workspace_bytes(self._h))

    PROF_SLOTS = {"ipa_core": 0, "edge_transition": 1, "edge_embed": 2, "ipa_total": 3, "seq_tfmr": 4, "forward": 5, "ipa_attn": 6}

    def profile_enable(self, on: bool = True):
        self._ck(lib().fdpt_profile_enable(self._h, int(on)))

    def profile_read(self) -> dict[str, tuple[int, float]]:
        """{slot: (launch count, total ms)} measured with CUDA events on the launching stream; clears the records."""
        out = {}
        for name, slot in self.PROF_SLOTS.items():
            n, ms = C.c_int(0), C.c_double(0.0)
            self._ck(lib().fdpt_profile_read(self._h, slot, C.byref(n), C.byref(ms)))
            out[name] = (n.value, ms.value)
        return out

    # ---- parameters
    def load_state_dict(self, sd: dict, strict: bool = True):
        expected = {k for k, _, _ in param_specs(self.dims, self.with_aatype)}
        unexpected = [k for k in sd if k not in expected]
        if strict and unexpected:
            raise RuntimeError(f"Unexpected key(s) in state_dict: {unexpected[:5]}")
        for k, v in sd.items():
            if k not in expected:
                continue
            t = torch.as_tensor(v).detach().to("cpu", torch.float32).contiguous()
            shape = (C.c_int64 * t.dim())(*t.shape)
            self._ck(lib().fdpt_load_param(self._h, k.encode(), t.data_ptr(), shape, t.dim()))
        self._ck(lib().fdpt_finalize_params(self._h))
        self._params_loaded = True

    def reserve(self, B: int, N: int):
        self._ck(lib().fdpt_reserve(self._h, B, N))

    # ---- forward
    def set_score_table(self, score_norms: np.ndarray | None, omega_bounds: np.ndarray | None = None):
        """Installs (or with None removes) the cached IGSO(3) score-norm table of so3.use_cached_score=True."""
        if score_norms is None:
            self._ck(lib().fdpt_set_score_table(self._h, None, 0, 0, None))
            return
        tab = np.ascontiguousarray(score_norms, np.float64)
        bnd = np.ascontiguousarray(omega_bounds, np.float64)
        assert tab.ndim == 2 and bnd.shape == (tab.shape[1] - 1,)
        self._ck(lib().fdpt_set_score_table(self._h, tab.ctypes.data, tab.shape[0], tab.shape[1], bnd.ctypes.data))

    def forward(self, pf: PreparedFeats, t: torch.Tensor, sigma: np.ndarray, want_backbone: bool = True,
                sigma_idx: np.ndarray | None = None) -> dict[str, torch.Tensor]:
        B, N, dev = pf.B, pf.N, self.device
        t32 = torch.as_tensor(t).to("cpu", torch.float32).reshape(-1)
        if t32.numel() != B:
            raise ValueError(f"t should have shape ({B},), got {tuple(t32.shape)}")
        pf.t_emb = timestep_embedding(t32, self.dims.index_embed_size).to(dev).contiguous()
        pf.t32 = t32.to(dev)
        pf.sigma = torch.as_tensor(np.asarray(sigma, np.float64).reshape(B)).to(dev)
        pf.sigma_idx = None if sigma_idx is None else torch.as_tensor(np.asarray(sigma_idx).reshape(B).astype(np.int32)).to(dev)
        out = {
            "rigids": torch.empty(B, N, 7, device=dev), "rot_score": torch.empty(B, N, 3, device=dev, dtype=torch.float64),
            "trans_score": torch.empty(B, N, 3, device=dev), "psi": torch.empty(B, N, 2, device=dev),
        }
        bb = torch.empty(B, N, 5, 3, device=dev) if want_backbone else None
        o = _Out(_ptr(out["rigids"]), _ptr(out["rot_score"]), _ptr(out["trans_score"]), _ptr(out["psi"]), _ptr(bb))
        fs = pf.struct()
        self._ck(lib().fdpt_forward(self._h, B, N, C.byref(fs), C.byref(o), self.stream))
        if bb is not None:
            out["atom37_bb"] = bb
        return out

    # ---- sampling loop
    def sample(self, pf: PreparedFeats, sched: np.ndarray, t_emb_tab: torch.Tensor, noise: torch.Tensor | None, self_condition=True,
               center=True, diffuse_rot=True, diffuse_trans=True, final_only=False, out: dict | None = None,
               philox_seed: int = 0, progress_chunk: int = 0) -> dict[str, torch.Tensor]:
        """Enqueue the whole reverse-diffusion loop (fdpt_sample).  `out` may hold pre-allocated trajectory buffers from
        `alloc_traj` (a fresh cudaMalloc inside torch.empty can take tens of milliseconds; callers that time the loop allocate first).
        noise=None: throughput mode, normals drawn on the device (Philox, `philox_seed`).  progress_chunk > 0: an event is recorded
        every that many steps so that `wait_step` lets the caller read finished trajectory slots while later steps run."""
        import time as _time

        _t0 = _time.perf_counter()
        B, N, dev = pf.B, pf.N, self.device
        T = sched.shape[0]
        sched = np.ascontiguousarray(sched, np.float64)
        assert sched.shape == (T, SCHED_COLS)
        if out is None:
            out = self.alloc_traj(B, N, T, final_only)
        else:
            ref = self.alloc_shapes(B, N, T, final_only)
            assert all(tuple(out[k].shape) == ref[k] and out[k].is_cuda and out[k].dtype == torch.float32 for k in ref), "out buffers do not match"
        _t1 = _time.perf_counter()
        tr = _Traj(_ptr(out["prot_traj"]), _ptr(out["rigid_traj"]), _ptr(out["trans_traj"]), _ptr(out["rigid_0_traj"]),
                   _ptr(out["psi_pred"]), int(final_only))
        t_emb_tab = t_emb_tab.to(dev, torch.float32).contiguous()
        if noise is not None:
            n_rev = int((sched[:, 7] == 0).sum())
            assert noise.dtype == torch.float64 and noise.is_cuda and noise.shape[0] >= max(n_rev, T - 1) and \
                tuple(noise.shape[1:]) == (2, B, N, 3), (noise.shape, noise.dtype)
        fs = pf.struct()
        _t2 = _time.perf_counter()
        self._ck(lib().fdpt_set_progress_chunk(self._h, int(progress_chunk)))
        self._ck(lib().fdpt_sample(self._h, B, N, C.byref(fs), T, sched.ctypes.data_as(C.POINTER(C.c_double)), _ptr(t_emb_tab),
                                   _ptr(noise), int(philox_seed) & 0xFFFFFFFFFFFFFFFF, int(self_condition), int(center), int(diffuse_rot),
                                   int(diffuse_trans), C.byref(tr), self.stream))
        _t3 = _time.perf_counter()
        self.last_sample_host_ms = {"alloc": (_t1 - _t0) * 1e3, "prep": (_t2 - _t1) * 1e3, "c_call": (_t3 - _t2) * 1e3}
        out["_keepalive"] = (t_emb_tab, noise, sched)
        return out

    def wait_step(self, step: int):
        """Blocks until timestep `step` of the most recent `sample(..., progress_chunk=c)` call has completed on the device."""
        self._ck(lib().fdpt_wait_step(self._h, int(step)))

    @staticmethod
    def alloc_shapes(B, N, T, final_only=False) -> dict:
        Ts = 1 if final_only else T
        return {"prot_traj": (Ts, B, N, 5, 3), "rigid_traj": (Ts if final_only else T + 1, B, N, 7), "trans_traj": (Ts, B, N, 3),
                "rigid_0_traj": (Ts, B, N, 5, 3), "psi_pred": (B, N, 2)}

    def alloc_traj(self, B, N, T, final_only=False) -> dict[str, torch.Tensor]:
        return {k: torch.empty(*shp, device=self.device) for k, shp in self.alloc_shapes(B, N, T, final_only).items()}

    # ---- unit entry points
    def linear(self, x, w, b, act=0):
        M, K = x.shape
        Nn = w.shape[0]
        y = torch.empty(M, Nn, device=self.device)
        self._ck(lib().fdpt_linear(self._h, M, Nn, K, _ptr(x), _ptr(w), _ptr(b), act, _ptr(y), self.stream))
        return y

    def set_option(self, option: int, value: int):
        self._ck(lib().fdpt_set_option(self._h, option, value))

    def bench_linear(self, x, w, b, reps=20) -> float:
        """us per launch of the Linear kernel on x[M,K] @ w[N,K]^T (weights split once)."""
        M, K = x.shape
        Nn = w.shape[0]
        y = torch.empty(M, Nn, device=self.device)
        ms = C.c_float(0)
        self._ck(lib().fdpt_bench_linear(self._h, M, Nn, K, _ptr(x), _ptr(w), _ptr(b), _ptr(y), reps, C.byref(ms)))
        return ms.value * 1e3

    def debug_read(self, n: int = 8 * 48) -> np.ndarray:
        buf = (C.c_int64 * n)()
        lib().fdpt_debug_read(self._h, buf, n)
        return np.array(buf[:], np.int64)

    def matmul(self, a, b, b_kmajor=True, alpha=1.0):
        """Batched a[Bt,M,K] @ (b[Bt,N,K]^T if b_kmajor else b[Bt,K,N]) on the node-side GEMM kernel."""
        Bt, M, K = a.shape
        Nn = b.shape[1] if b_kmajor else b.shape[2]
        c = torch.empty(Bt, M, Nn, device=self.device)
        self._ck(lib().fdpt_matmul(self._h, Bt, M, Nn, K, _ptr(a), a.stride(1), a.stride(0), _ptr(b), b.stride(1), b.stride(0),
                                   int(b_kmajor), alpha, _ptr(c), Nn, M * Nn, self.stream))
        return c

    def tc_linear(self, x, w, b, act=0):
        M, K = x.shape
        Nn = w.shape[0]
        y = torch.empty(M, Nn, device=self.device)
        self._ck(lib().fdpt_tc_linear(self._h, M, Nn, K, _ptr(x), _ptr(w), _ptr(b), act, _ptr(y), self.stream))
        return y

    def ipa(self, blk, s, z, quats, trans, mask):
        B, N, _ = s.shape
        out = torch.empty(B, N, self.dims.c_s, device=self.device)
        self._ck(lib().fdpt_ipa(self._h, blk, B, N, _ptr(s), _ptr(z), _ptr(quats), _ptr(trans), _ptr(mask), _ptr(out), self.stream))
        return out

    def edge_transition(self, blk, node, z, mask):
        B, N, _ = node.shape
        out = torch.empty_like(z)
        self._ck(lib().fdpt_edge_transition(self._h, blk, B, N, _ptr(node), _ptr(z), _ptr(mask), _ptr(out), self.stream))
        return out

    def embed(self, pf: PreparedFeats, t: torch.Tensor):
        B, N = pf.B, pf.N
        t32 = torch.as_tensor(t).to("cpu", torch.float32).reshape(-1)
        pf.t_emb = timestep_embedding(t32, self.dims.index_embed_size).to(self.device).contiguous()
        pf.t32 = t32.to(self.device)
        pf.sigma = torch.zeros(B, dtype=torch.float64, device=self.device)
        node = torch.empty(B, N, self.dims.c_s, device=self.device)
        edge = torch.empty(B, N, N, self.dims.c_z, device=self.device)
        fs = pf.struct()
        self._ck(lib().fdpt_embed(self._h, B, N, C.byref(fs), _ptr(node), _ptr(edge), self.stream))
        return node, edge

    def reverse(self, rigids_t, rot_score, trans_score, dmask, z_rot, z_trans, sched_row, center=True, diffuse_rot=True, diffuse_trans=True):
        B, N, _ = rigids_t.shape
        out = torch.empty_like(rigids_t)
        row = np.ascontiguousarray(sched_row, np.float64)
        self._ck(lib().fdpt_reverse(self._h, B, N, _ptr(rigids_t), _ptr(rot_score), _ptr(trans_score), _ptr(dmask), _ptr(z_rot), _ptr(z_trans),
                                    row.ctypes.data_as(C.POINTER(C.c_double)), int(center), int(diffuse_rot), int(diffuse_trans), _ptr(out),
                                    self.stream))
        torch.cuda.current_stream(self.device).synchronize()  # `row` is pageable host memory
        return out

    def backbone(self, rigids, psi, aatype):
        B, N, _ = rigids.shape
        out = torch.empty(B, N, 5, 3, device=self.device)
        self._ck(lib().fdpt_backbone(self._h, B, N, _ptr(rigids), _ptr(psi), _ptr(aatype), _ptr(out), self.stream))
        return out

    def rot_score_idx(self, quats_t, quats_0, sigma_idx, mask=None):
        """Rotation score by look-up in the installed score table (so3.use_cached_score=True)."""
        B, N, _ = quats_t.shape
        out = torch.empty(B, N, 3, device=self.device, dtype=torch.float64)
        self._ck(lib().fdpt_rot_score_idx(self._h, B, N, _ptr(quats_t), _ptr(quats_0), _ptr(sigma_idx), _ptr(mask), _ptr(out), self.stream))
        return out

    def sample_ref(self, B, N, impute, diffuse_mask, cdf, omega_grid, draws=None, philox_seed=0, diffuse_rot=True, diffuse_trans=True):
        """x_T of B samples of one structure on the device (fdpt_sample_ref).  impute [N,7] / diffuse_mask [N] float32 or None;
        cdf / omega_grid float64 [num_omega]; draws float64 [B,7N] (parity mode) or None (Philox)."""
        out = torch.empty(B, N, 7, device=self.device)
        self._ck(lib().fdpt_sample_ref(self._h, B, N, _ptr(impute), _ptr(diffuse_mask), _ptr(cdf), _ptr(omega_grid), int(cdf.numel()),
                                       _ptr(draws), int(philox_seed) & 0xFFFFFFFFFFFFFFFF, int(diffuse_rot), int(diffuse_trans), _ptr(out),
                                       self.stream))
        return out

    def tmem_a_selftest(self, a, b):
        """d = fp16(a[128,64]) @ fp16(b[128,64])^T through tcgen05.mma with A in tensor memory (bring-up unit)."""
        d = torch.empty(128, 128, device=self.device)
        self._ck(lib().fdpt_tmem_a_selftest(self._h, _ptr(a), _ptr(b), _ptr(d), self.stream))
        return d

    def seq_tfmr(self, blk, node, node0, mask):
        """(encoder output [B,N,320], node + post_tfmr(...) [B,N,256]) of block `blk`'s sequence-transformer sub-block."""
        B, N, _ = node.shape
        tf = torch.empty(B, N, self.dims.c_s + self.dims.c_skip, device=self.device)
        out = torch.empty(B, N, self.dims.c_s, device=self.device)
        self._ck(lib().fdpt_seq_tfmr(self._h, blk, B, N, _ptr(node), _ptr(node0), _ptr(mask), _ptr(tf), _ptr(out), self.stream))
        return tf, out

    def rot_score(self, quats_t, quats_0, sigma, mask=None):
        B, N, _ = quats_t.shape
        out = torch.empty(B, N, 3, device=self.device, dtype=torch.float64)
        self._ck(lib().fdpt_rot_score(self._h, B, N, _ptr(quats_t), _ptr(quats_0), _ptr(sigma), _ptr(mask), _ptr(out), self.stream))
        return out

    def trans_score(self, trans_t, trans_0, t32, mask=None, scale=True):
        B, N, _ = trans_t.shape
        out = torch.empty(B, N, 3, device=self.device)
        self._ck(lib().fdpt_trans_score(self._h, B, N, _ptr(trans_t), _ptr(trans_0), _ptr(t32), _ptr(mask), int(scale), _ptr(out), self.stream))
        return out


# --------------------------------------------------------------------------------------------------------
# device-executed SE3Diffuser methods (bound from se3_diffuser.py)
# --------------------------------------------------------------------------------------------------------
_default_ctx: dict[tuple, Context] = {}


def default_context(device: int | None = None, diffuser=None) -> Context:
    """Context for the standalone diffuser calls (reverse / calc_*_score), built from the diffuser's OWN r3 configuration
    (min_b, max_b, coordinate_scaling: r3_diffuser.py:15-35) on the current CUDA device."""
    if device is None:
        device = torch.cuda.current_device() if torch.cuda.is_available() else 0
    r3 = getattr(diffuser, "_r3_diffuser", None)
    key = (device, float(r3.min_b) if r3 else 0.1, float(r3.max_b) if r3 else 20.0, float(r3.coordinate_scaling) if r3 else 0.1)
    if key not in _default_ctx:
        _default_ctx[key] = Context(device=device, r3_min_b=key[1], r3_max_b=key[2], r3_coordinate_scaling=key[3])
    return _default_ctx[key]


def reverse_host_api(diffuser, rigid_t, rot_score, trans_score, t, dt, diffuse_mask, center, noise_scale):
    """SE3Diffuser.reverse with the reference's argument meaning (numpy scores in, Rigid out); noise drawn from the
    global legacy numpy RNG in the reference's order (rot, then trans)."""
    from .rigid import Rigid

    ctx = default_context(diffuser=diffuser)
    dev = ctx.device
    r7 = rigid_t.to_tensor_7().to(dev, torch.float32)
    squeeze = r7.dim() == 2
    if squeeze:
        r7 = r7[None]
    B, N, _ = r7.shape
    rs = torch.as_tensor(np.asarray(rot_score, np.float64)).reshape(B, N, 3).to(dev)
    ts = torch.as_tensor(np.asarray(trans_score, np.float32)).reshape(B, N, 3).to(dev)
    z_rot = torch.as_tensor(np.random.normal(size=(B, N, 3))).to(dev) if diffuser._diffuse_rot else torch.zeros(B, N, 3, dtype=torch.float64, device=dev)
    z_tr = torch.as_tensor(np.random.normal(size=(B, N, 3))).to(dev) if diffuser._diffuse_trans else torch.zeros(B, N, 3, dtype=torch.float64, device=dev)
    dm = np.ones((B, N), np.float32) if diffuse_mask is None else np.asarray(diffuse_mask, np.float32).reshape(B, N)
    row = diffuser.step_scalars(float(t), float(dt), float(noise_scale))
    out = ctx.reverse(r7.contiguous(), rs.contiguous(), ts.contiguous(), torch.as_tensor(dm).to(dev), z_rot, z_tr, row, center,
                      diffuser._diffuse_rot, diffuser._diffuse_trans)
    if squeeze:
        out = out[0]
    return Rigid.from_tensor_7(out.cpu())


def rot_score_host_api(diffuser, rots_t, rots_0, t):
    ctx = default_context(diffuser=diffuser)
    dev = ctx.device
    qt = rots_t.get_quats().to(dev, torch.float32).contiguous()
    q0 = rots_0.get_quats().to(dev, torch.float32).contiguous()
    t_np = torch.as_tensor(t).detach().cpu().numpy()
    so3 = diffuser._so3_diffuser
    if so3.use_cached_score:
        if getattr(ctx, "_table_of", None) is not so3:
            ctx.set_score_table(so3.score_norms, so3.discrete_omega[:-1])
            ctx._table_of = so3
        idx = torch.as_tensor(np.asarray(so3.t_to_idx(t_np)).reshape(-1).astype(np.int32)).to(dev)
        return ctx.rot_score_idx(qt, q0, idx)
    if getattr(ctx, "_table_of", None) is not None:
        ctx.set_score_table(None)
        ctx._table_of = None
    sigma = torch.as_tensor(np.asarray(so3.grid_sigma(t_np), np.float64).reshape(-1)).to(dev)
    return ctx.rot_score(qt, q0, sigma)


def trans_score_host_api(diffuser, trans_t, trans_0, t, scale):
    ctx = default_context(diffuser=diffuser)
    dev = ctx.device
    xt = torch.as_tensor(trans_t).to(dev, torch.float32).contiguous()
    x0 = torch.as_tensor(trans_0).to(dev, torch.float32).contiguous()
    t32 = torch.as_tensor(t).to(dev, torch.float32).reshape(-1).contiguous()
    return ctx.trans_score(xt, x0, t32, None, scale)
