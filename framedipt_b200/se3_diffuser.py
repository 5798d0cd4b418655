"""Host-side mirror of the reference's ``SE3Diffuser`` call surface.

Same constructor argument (an attribute-style config with ``.so3.*``, ``.r3.*``, ``.diffuse_rot``,
``.diffuse_trans``), same method names and argument meaning as
framedipt/diffusion/se3_diffuser.py:39-529.  What differs is *where* things run:

* ``sample_ref`` (x_T, once per sample, legacy global numpy RNG in the reference's draw order) stays on
  the host — it is RNG bookkeeping, not the hot path.
* ``reverse`` / ``calc_rot_score`` / ``calc_trans_score`` are executed by the sm_100a kernels in
  ``libfdpt.so`` (no CPU fallback: they raise if the CUDA library is unavailable).
* per-timestep scalars (sigma(t) on the 1000-point grid, g_so3, b_t ...) are computed here with the same
  numpy expressions as the reference (so3_diffuser.py:288-323, r3_diffuser.py:48-96) and shipped to the
  device as a schedule table, so the loop never synchronises with the host.
"""
from __future__ import annotations

import numpy as np
import torch

from .rigid import Rigid, Rotation


def _get(conf, name, default=None):
    try:
        return getattr(conf, name)
    except (AttributeError, KeyError):
        return conf.get(name, default) if hasattr(conf, "get") else default


class SO3Schedule:
    """sigma / diffusion-coefficient schedule of IGSO(3) (so3_diffuser.py:288-323)."""

    def __init__(self, so3_conf):
        self.schedule = _get(so3_conf, "schedule", "logarithmic")
        if self.schedule != "logarithmic":
            raise ValueError(f"Unrecognize schedule {self.schedule}")
        self.min_sigma = float(_get(so3_conf, "min_sigma", 0.1))
        self.max_sigma = float(_get(so3_conf, "max_sigma", 1.5))
        self.num_sigma = int(_get(so3_conf, "num_sigma", 1000))
        self.num_omega = int(_get(so3_conf, "num_omega", 1000))
        self.use_cached_score = bool(_get(so3_conf, "use_cached_score", False))
        if self.use_cached_score:
            raise NotImplementedError("use_cached_score=True (table lookup score) is out of scope; the series is evaluated on device")
        self.discrete_omega = np.linspace(0, np.pi, self.num_omega + 1)[1:]
        self._cdf_rows: dict[int, np.ndarray] = {}

    def sigma(self, t):
        t = np.asarray(t)
        if np.any(t < 0) or np.any(t > 1):
            raise ValueError(f"Invalid t={t}")
        return np.log(t * np.exp(self.max_sigma) + (1 - t) * np.exp(self.min_sigma))

    @property
    def discrete_sigma(self):
        return self.sigma(np.linspace(0.0, 1.0, self.num_sigma))

    def t_to_idx(self, t):
        return np.digitize(self.sigma(t), self.discrete_sigma) - 1

    def grid_sigma(self, t):
        return self.discrete_sigma[self.t_to_idx(t)]

    def diffusion_coef(self, t):
        s = self.sigma(t)
        return np.sqrt(2 * (np.exp(self.max_sigma) - np.exp(self.min_sigma)) * s / np.exp(s))

    def _series(self, sig):
        lv = np.arange(1000)[None]
        om = self.discrete_omega[:, None]
        return ((2 * lv + 1) * np.exp(-lv * (lv + 1) * sig ** 2 / 2) * np.sin(om * (lv + 0.5)) / np.sin(om / 2)).sum(-1)

    def cdf_row(self, idx: int) -> np.ndarray:
        """One row of the reference's ``_cdf`` table (so3_diffuser.py:247-262), built lazily."""
        if idx not in self._cdf_rows:
            pdf = self._series(self.discrete_sigma[idx]) * (1 - np.cos(self.discrete_omega)) / np.pi
            self._cdf_rows[idx] = pdf.cumsum() / self.num_omega * np.pi
        return self._cdf_rows[idx]

    def score_scaling(self, t):
        """so3_diffuser.py:280-285, 404-406 for a single t (row built lazily)."""
        idx = int(self.t_to_idx(t))
        sig = self.discrete_sigma[idx]
        lv = np.arange(1000)[None]
        om = self.discrete_omega[:, None]
        expn = self._series(sig)
        hi, dhi = np.sin(om * (lv + 0.5)), (lv + 0.5) * np.cos(om * (lv + 0.5))
        lo, dlo = np.sin(om / 2), 0.5 * np.cos(om / 2)
        ds = ((2 * lv + 1) * np.exp(-lv * (lv + 1) * sig ** 2 / 2) * (lo * dhi - hi * dlo) / lo ** 2).sum(-1)
        norms = ds / (expn + 1e-4)
        pdf = expn * (1 - np.cos(self.discrete_omega)) / np.pi
        return np.sqrt(np.abs(np.sum(norms ** 2 * pdf) / np.sum(pdf))) / np.sqrt(3)

    def sample(self, t: float, n_samples: int = 1) -> np.ndarray:
        """so3_diffuser.py:325-357 (global legacy numpy RNG: randn(n,3) then rand(n))."""
        x = np.random.randn(n_samples, 3)
        x /= np.linalg.norm(x, axis=-1, keepdims=True)
        u = np.random.rand(n_samples)
        ang = np.interp(u, self.cdf_row(int(self.t_to_idx(t))), self.discrete_omega)
        return x * ang[:, None]


class R3Schedule:
    """VP-SDE schedule on translations (r3_diffuser.py:12-96, 387-408)."""

    def __init__(self, r3_conf):
        self.min_b = float(_get(r3_conf, "min_b", 0.1))
        self.max_b = float(_get(r3_conf, "max_b", 20.0))
        self.coordinate_scaling = float(_get(r3_conf, "coordinate_scaling", 0.1))

    def b_t(self, t):
        t = np.asarray(t)
        if np.any(t < 0) or np.any(t > 1):
            raise ValueError(f"Invalid t={t}")
        return self.min_b + t * (self.max_b - self.min_b)

    def marginal_b_t(self, t):
        return t * self.min_b + 0.5 * (t ** 2) * (self.max_b - self.min_b)

    def conditional_var(self, t):
        return 1 - np.exp(-self.marginal_b_t(t))

    def score_scaling(self, t):
        return 1 / np.sqrt(self.conditional_var(t))


class SE3Diffuser:
    def __init__(self, se3_conf) -> None:
        self._se3_conf = se3_conf
        self._diffuse_rot = bool(_get(se3_conf, "diffuse_rot", True))
        self._diffuse_trans = bool(_get(se3_conf, "diffuse_trans", True))
        self._so3_diffuser = SO3Schedule(se3_conf.so3)
        self._r3_diffuser = R3Schedule(se3_conf.r3)
        # the reference seeds the global numpy RNG in both sub-diffuser constructors
        # (so3_diffuser.py:286, r3_diffuser.py:24)
        for sub in (se3_conf.so3, se3_conf.r3):
            seed = _get(sub, "seed", None)
            np.random.seed(seed)

    # ---- schedule -------------------------------------------------------------------------------
    def score_scaling(self, t: float):
        return self._so3_diffuser.score_scaling(t), self._r3_diffuser.score_scaling(t)

    def step_scalars(self, t: float, dt: float, noise_scale: float) -> np.ndarray:
        """Per-step scalar row consumed by the device schedule table (see include/fdpt.h, FDPT_SCHED_*).
        t is rounded through float32 where the reference does (feats["t"] is a float32 tensor,
        experiments/utils.py:186; scores use it, the reverse step uses the float64 t)."""
        t32 = float(np.float32(t))
        so3, r3 = self._so3_diffuser, self._r3_diffuser
        sigma = float(so3.grid_sigma(np.array(np.float32(t))))  # float64 arithmetic on the fp32-rounded t
        g = float(so3.diffusion_coef(t))
        b_t = float(r3.b_t(t))
        t32f = np.float32(t)
        mb = np.float32(t32f * np.float32(r3.min_b) + np.float32(0.5) * (t32f * t32f) * np.float32(r3.max_b - r3.min_b))
        return np.array([t32, sigma, g * g * dt, g * np.sqrt(dt) * noise_scale, b_t, dt, np.sqrt(b_t) * np.sqrt(dt) * noise_scale,
                         float(mb)], dtype=np.float64)

    # ---- x_T ------------------------------------------------------------------------------------
    def sample_ref(self, n_samples: int, chain_index=None, impute: Rigid | None = None, diffuse_mask=None,
                   as_tensor_7: bool = False):
        """se3_diffuser.py:455-529 (host numpy; legacy global RNG, draw order randn(n,3), rand(n), normal([n_diff,3]))."""
        from scipy.spatial.transform import Rotation as SR

        if impute is None:
            if not self._diffuse_rot:
                raise ValueError("Must provide impute values as we're not diffusing rotations!")
            if not self._diffuse_trans:
                raise ValueError("Must provide impute values as we're not diffusing translations!")
            if diffuse_mask is not None:
                raise ValueError("Must provide imputation values for unmasked regions!")
            trans_impute = np.zeros((n_samples, 3), np.float32)
            rot_impute = np.zeros((n_samples, 3))
        else:
            if impute.shape[0] != n_samples:
                raise ValueError(f"impute should have shape ({n_samples}, ...), got {impute.shape}.")
            R = impute.get_rots().get_rot_mats().cpu().numpy().reshape(-1, 3, 3)
            rot_impute = SR.from_matrix(R).as_rotvec().reshape(n_samples, 3)
            trans_impute = impute.get_trans().cpu().numpy().reshape(n_samples, 3)
        if diffuse_mask is not None:
            diffuse_mask = np.asarray(diffuse_mask)
        rot_ref = self._so3_diffuser.sample(1.0, n_samples) if self._diffuse_rot else rot_impute
        if self._diffuse_trans:
            cs = self._r3_diffuser.coordinate_scaling
            x_ref = trans_impute * cs
            bm = diffuse_mask.astype(bool) if diffuse_mask is not None else np.ones(n_samples, bool)
            loc = np.zeros_like(trans_impute[bm])
            inp = np.random.normal(loc=loc, scale=np.ones_like(loc))
            x_out = x_ref.copy()
            x_out[bm] = inp
            trans_ref = x_out / cs
        else:
            trans_ref = trans_impute
        if diffuse_mask is not None:
            dm = diffuse_mask[..., None]
            rot_ref = dm * rot_ref + (1 - dm) * rot_impute
        rotmat = SR.from_rotvec(rot_ref.reshape(-1, 3)).as_matrix().reshape(n_samples, 3, 3)
        rigids_t = Rigid(Rotation(rot_mats=torch.Tensor(rotmat)), torch.tensor(trans_ref))
        if as_tensor_7:
            rigids_t = rigids_t.to_tensor_7()
        return {"rigids_t": rigids_t}

    # ---- device-executed pieces (bound lazily to avoid importing the CUDA library for host-only use) ----
    def reverse(self, rigid_t: Rigid, rot_score, trans_score, t: float, dt: float, diffuse_mask=None,
                chain_indices=None, center: bool = True, noise_scale: float = 1.0) -> Rigid:
        from . import runtime

        return runtime.reverse_host_api(self, rigid_t, rot_score, trans_score, t, dt, diffuse_mask, center, noise_scale)

    def calc_rot_score(self, rots_t: Rotation, rots_0: Rotation, t: torch.Tensor) -> torch.Tensor:
        from . import runtime

        return runtime.rot_score_host_api(self, rots_t, rots_0, t)

    def calc_trans_score(self, trans_t, trans_0, t, use_torch: bool = False, scale: bool = True):
        from . import runtime

        return runtime.trans_score_host_api(self, trans_t, trans_0, t, scale)
