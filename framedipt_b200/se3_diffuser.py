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

from .rigid import Rigid, Rotation, rotmats_to_rigid


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
        self.cache_dir = _get(so3_conf, "cache_dir", None)
        self.discrete_omega = np.linspace(0, np.pi, self.num_omega + 1)[1:]
        self._cdf_rows: dict[int, np.ndarray] = {}
        self._score_norms: np.ndarray | None = None

    def _cache_path(self) -> str | None:
        """Directory of the reference's IGSO(3) cache for this configuration (so3_diffuser.py:207-233): the same files are read
        and written, so a cache built by either implementation serves both."""
        import os

        if not self.cache_dir:
            return None
        rp = lambda x: str(x).replace(".", "_")  # noqa: E731
        return os.path.join(self.cache_dir, f"eps_{self.num_sigma}_omega_{self.num_omega}_min_sigma_{rp(self.min_sigma)}_"
                                            f"max_sigma_{rp(self.max_sigma)}_schedule_{self.schedule}")

    @property
    def score_norms(self) -> np.ndarray:
        """The reference's `_score_norms` table [num_sigma, num_omega] (so3_diffuser.py:264-273: `score(exp_vals[i], discrete_omega,
        sigma_i)` per grid sigma), needed only for so3.use_cached_score=True.  Loaded from the reference's cache file when present,
        else computed once (float64 numpy, 1000-term series; about a minute) and saved there."""
        import os

        if self._score_norms is None:
            d = self._cache_path()
            f = os.path.join(d, "score_norms.npy") if d else None
            if f and os.path.exists(f):
                self._score_norms = np.load(f)
            else:
                self._score_norms = np.asarray([self.score_norm_row(i) for i in range(self.num_sigma)])
                if f:
                    try:
                        os.makedirs(d, exist_ok=True)
                        np.save(f, self._score_norms)
                    except OSError:
                        pass
        return self._score_norms

    def score_norm_row(self, idx: int) -> np.ndarray:
        """Row `idx` of `_score_norms`: d/d omega log f(omega; sigma_idx) on the omega grid (so3_diffuser.py:122-191, 264-273)."""
        if not hasattr(self, "_series_basis"):
            lv = np.arange(1000)[None]
            om = self.discrete_omega[:, None]
            hi, dhi = np.sin(om * (lv + 0.5)), (lv + 0.5) * np.cos(om * (lv + 0.5))
            lo, dlo = np.sin(om / 2), 0.5 * np.cos(om / 2)
            self._series_basis = ((lo * dhi - hi * dlo) / lo ** 2, hi / lo)
        base, ser = self._series_basis
        lv = np.arange(1000)[None]
        sig = self.discrete_sigma[idx]
        c = (2 * lv + 1) * np.exp(-lv * (lv + 1) * sig ** 2 / 2)
        return (c * base).sum(-1) / ((c * ser).sum(-1) + 1e-4)

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


    # ---- one-step forward noising and the two transition log-densities (EigenFold confidence score, so3_diffuser.py:408-567)
    def forward(self, x_t_1, t_1: float, dt: float, diffuse_mask=None, chain_indices=None, noise_scale: float = 1.0):
        z = noise_scale * np.random.normal(size=x_t_1.shape)
        step = self.diffusion_coef(t_1) * np.sqrt(dt) * z
        if diffuse_mask is not None:
            step = step * diffuse_mask[..., None]
        return _compose_rotvec(x_t_1.reshape(-1, 3), step.reshape(-1, 3)).reshape(x_t_1.shape)

    def distribution(self, rot_t, score_t, t: float, dt: float, diffuse_mask=None, chain_indices=None):
        g = self.diffusion_coef(t)
        drift = (g ** 2) * score_t * dt
        if diffuse_mask is not None:
            drift = drift * diffuse_mask[..., None]
        mu = _compose_rotvec(rot_t.reshape(-1, 3), drift.reshape(-1, 3)).reshape(rot_t.shape)
        return mu, g * np.sqrt(dt)

    def log_prob_forward(self, rot_t, rot_t_1, t_1: float, dt: float, diffuse_mask=None, chain_indices=None) -> float:
        std = self.diffusion_coef(t_1) * np.sqrt(dt)
        return gaussian_log_prob(rot_t_1, std, align_rotation_vectors(rot_t, rot_t_1), diffuse_mask)

    def log_prob_backward(self, rot_t, rot_t_1, score_t, t: float, dt: float, diffuse_mask=None, chain_indices=None) -> float:
        mu, std = self.distribution(rot_t, score_t, t, dt, diffuse_mask)
        return gaussian_log_prob(mu, std, align_rotation_vectors(rot_t_1, mu), diffuse_mask)


def _rotvec_of(rigid: Rigid):
    """(trans [.., 3] fp32, rotation vectors [.., 3] fp64) of a Rigid, the representation every diffuser step works in
    (se3_diffuser.py:16-23: fp32 rotation matrices -> scipy rotvec)."""
    from scipy.spatial.transform import Rotation as SR

    R = rigid.get_rots().get_rot_mats().cpu().numpy()
    rv = SR.from_matrix(R.reshape(-1, 3, 3)).as_rotvec().reshape(R.shape[:-2] + (3,))
    return rigid.get_trans().cpu().numpy(), rv


def _rigid_of(rotvec: np.ndarray, trans: np.ndarray) -> Rigid:
    """Inverse of _rotvec_of; rotation matrices and translations are stored in fp32 (se3_diffuser.py:26-36)."""
    from scipy.spatial.transform import Rotation as SR

    mats = SR.from_rotvec(rotvec.reshape(-1, 3)).as_matrix().reshape(rotvec.shape[:-1] + (3, 3))
    return rotmats_to_rigid(mats, trans)


def _compose_rotvec(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """rotvec(R(a) R(b)) (right multiplication, data/transforms.py:33-38)."""
    from scipy.spatial.transform import Rotation as SR

    Ra, Rb = SR.from_rotvec(a.reshape(-1, 3)).as_matrix(), SR.from_rotvec(b.reshape(-1, 3)).as_matrix()
    return SR.from_matrix(np.einsum("nij,njk->nik", Ra, Rb)).as_rotvec().reshape(a.shape)


def align_rotation_vectors(inputs: np.ndarray, targets: np.ndarray) -> np.ndarray:
    """The representative of `inputs` (angle theta or 2 pi - theta about the flipped axis) whose axis points into the half space
    of `targets` (so3_diffuser.py:99-119)."""
    ang = np.linalg.norm(inputs, axis=-1, keepdims=True)
    axis = inputs / ang
    t_axis = targets / np.linalg.norm(targets, axis=-1, keepdims=True)
    sgn = np.sign(np.einsum("...i,...i->...", t_axis, axis))[..., None]
    return (axis * sgn) * np.where(sgn > 0, ang, 2 * np.pi - ang)


def gaussian_log_prob(mu, std, x, diffuse_mask=None) -> float:
    """Sum over the masked residues of log N(x; mu, std^2) (r3_utils.py:10-42).  Evaluated with torch tensors so that the arithmetic
    type follows the reference's: fp32 when both x and mu are fp32 positions (translation forward kernel; the scalar std adopts the
    tensor dtype as in torch.distributions.Normal), fp64 as soon as a score or rotation vector is involved; the masked values are
    summed by numpy (pairwise) in that dtype."""
    import math

    mu_t = torch.from_numpy(mu) if isinstance(mu, np.ndarray) else torch.as_tensor(mu)
    x_t = torch.from_numpy(x) if isinstance(x, np.ndarray) else torch.as_tensor(x)
    scale = torch.from_numpy(std) if isinstance(std, np.ndarray) else torch.as_tensor(float(std), dtype=mu_t.dtype)
    lp = -((x_t - mu_t) ** 2) / (2 * scale ** 2) - scale.log() - math.log(math.sqrt(2 * math.pi))
    if diffuse_mask is not None:
        sel = torch.as_tensor(np.asarray(diffuse_mask)).bool()[..., None]
        lp = torch.masked_select(lp, sel)
    return lp.cpu().numpy().sum()


class R3Schedule:
    """VP-SDE schedule on translations (r3_diffuser.py:12-96, 387-408)."""

    def __init__(self, r3_conf):
        self.min_b = float(_get(r3_conf, "min_b", 0.1))
        self.max_b = float(_get(r3_conf, "max_b", 20.0))
        self.coordinate_scaling = float(_get(r3_conf, "coordinate_scaling", 0.1))

    def b_t(self, t):
        if np.any(np.asarray(t) < 0) or np.any(np.asarray(t) > 1):
            raise ValueError(f"Invalid t={t}")
        return self.min_b + t * (self.max_b - self.min_b)  # keeps t's type (python float stays "weak" in numpy promotion)

    def diffusion_coef(self, t):
        return np.sqrt(self.b_t(t))

    def drift_coef(self, x, t):
        return -1 / 2 * self.b_t(t) * x

    def _scale(self, x):
        return x * self.coordinate_scaling

    def _unscale(self, x):
        return x / self.coordinate_scaling

    # ---- one-step forward noising and the two transition log-densities (EigenFold confidence score, r3_diffuser.py:122-260);
    #      expressions keep the reference's operand order so that numpy's type promotion (fp32 positions, fp64 noise) is the same
    def forward(self, x_t_1, t_1: float, dt: float, diffuse_mask=None, chain_indices=None, center: bool = True,
                noise_scale: float = 1.0):
        x_t_1 = self._scale(x_t_1)
        g_t = self.diffusion_coef(t_1)
        f_t = self.drift_coef(x_t_1, t_1)
        z = noise_scale * np.random.normal(size=x_t_1.shape)
        step = f_t * dt + g_t * np.sqrt(dt) * z
        if diffuse_mask is not None:
            step *= diffuse_mask[..., None]
        else:
            diffuse_mask = np.ones(x_t_1.shape[:-1])
        x_t = x_t_1 + step
        if center:
            com = np.sum(x_t, axis=-2) / np.sum(diffuse_mask, axis=-1)[..., None]
            x_t -= com[..., None, :]
        return self._unscale(x_t)

    def distribution(self, x_t, score_t, t: float, dt: float, diffuse_mask=None, chain_indices=None):
        x_t = self._scale(x_t)
        g_t = self.diffusion_coef(t)
        f_t = self.drift_coef(x_t, t)
        mu = x_t - (f_t - g_t ** 2 * score_t) * dt
        if diffuse_mask is not None:
            mu *= diffuse_mask[..., None]
        return mu, g_t * np.sqrt(dt)

    def log_prob_forward(self, x_t, x_t_1, t_1: float, dt: float, diffuse_mask, chain_indices=None) -> float:
        x_t_1 = self._scale(x_t_1)
        std = self.diffusion_coef(t_1) * np.sqrt(dt)
        mu = x_t_1 + self.drift_coef(x_t_1, t_1) * dt
        if diffuse_mask is not None:
            mu *= diffuse_mask[..., None]
        return gaussian_log_prob(mu, std, self._scale(x_t), diffuse_mask)

    def log_prob_backward(self, x_t, x_t_1, score_t, t: float, dt: float, diffuse_mask, chain_indices=None) -> float:
        if diffuse_mask is not None:
            diffuse_mask = diffuse_mask.astype(bool)
        mu, std = self.distribution(x_t, score_t, t, dt, diffuse_mask)
        return gaussian_log_prob(mu, std, self._scale(x_t_1), diffuse_mask)

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
        idx = float(so3.t_to_idx(np.array(np.float32(t))))  # row of the cached score table (use_cached_score)
        return np.array([t32, sigma, g * g * dt, g * np.sqrt(dt) * noise_scale, b_t, dt, np.sqrt(b_t) * np.sqrt(dt) * noise_scale,
                         float(mb), idx, 0.0], dtype=np.float64)

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

    # ---- EigenFold confidence score pieces (se3_diffuser.py:50-196): host numpy on [N, 3] arrays, as in the reference --------
    def forward(self, rigids_t_1: Rigid, t_1: float, dt: float, diffuse_mask=None, chain_indices=None) -> Rigid:
        """One forward-noising step x(t-1) -> x(t); draws the translation noise first, then the rotation noise, from the legacy
        global numpy RNG like the reference."""
        trans_0, rot_0 = _rotvec_of(rigids_t_1)
        trans_1 = self._r3_diffuser.forward(trans_0, t_1, dt, diffuse_mask, chain_indices, center=False)
        rot_1 = self._so3_diffuser.forward(rot_0, t_1, dt, diffuse_mask, chain_indices)
        if diffuse_mask is not None:
            dm = diffuse_mask[..., None]
            rot_1 = dm * rot_1 + (1 - dm) * rot_0
            trans_1 = dm * trans_1 + (1 - dm) * trans_0
        return _rigid_of(rot_1, trans_1)

    def log_prob_forward(self, rigids_t: Rigid, rigids_t_1: Rigid, t_1: float, dt: float, diffuse_mask=None, chain_indices=None) -> float:
        """log q(x(t) | x(t-1)) summed over the diffused residues."""
        trans_t, rot_t = _rotvec_of(rigids_t)
        trans_p, rot_p = _rotvec_of(rigids_t_1)
        return (self._r3_diffuser.log_prob_forward(trans_t, trans_p, t_1, dt, diffuse_mask)
                + self._so3_diffuser.log_prob_forward(rot_t, rot_p, t_1, dt, diffuse_mask))

    def log_prob_backward(self, rigids_t: Rigid, rigids_t_1: Rigid, trans_score_t, rot_score_t, t: float, dt: float, diffuse_mask,
                          chain_indices=None) -> float:
        """log p(x(t-1) | x(t)) of the learned reverse kernel, summed over the diffused residues."""
        trans_t, rot_t = _rotvec_of(rigids_t)
        trans_p, rot_p = _rotvec_of(rigids_t_1)
        return (self._r3_diffuser.log_prob_backward(trans_t, trans_p, trans_score_t, t, dt, diffuse_mask)
                + self._so3_diffuser.log_prob_backward(rot_t, rot_p, rot_score_t, t, dt, diffuse_mask))

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
