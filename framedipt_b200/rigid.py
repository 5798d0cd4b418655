"""Minimal host-side frame view types (`Rigid`, `Rotation`).

The reference passes frames around as ``openfold.utils.rigid_utils.Rigid`` objects
(openfold/utils/rigid_utils.py:853-1448).  The sampler call surface only needs a small part of
that API (``from_tensor_7``, ``to_tensor_7``, ``get_rots``, ``get_trans``, ``get_rot_mats``,
``get_quats``, ``shape``, ``identity``, indexing, ``.to``); this is a fresh, small implementation
of exactly that subset as a plain container of torch tensors.  Heavy math lives in the CUDA library.
"""
from __future__ import annotations

import numpy as np
import torch


def _quat_to_rot(q: torch.Tensor) -> torch.Tensor:
    a, b, c, d = q.unbind(-1)
    rows = [
        torch.stack([a * a + b * b - c * c - d * d, 2 * (b * c - a * d), 2 * (b * d + a * c)], -1),
        torch.stack([2 * (b * c + a * d), a * a - b * b + c * c - d * d, 2 * (c * d - a * b)], -1),
        torch.stack([2 * (b * d - a * c), 2 * (c * d + a * b), a * a - b * b - c * c + d * d], -1),
    ]
    return torch.stack(rows, -2)


def _rot_to_quat(R: torch.Tensor) -> torch.Tensor:
    """Unit quaternion (w,x,y,z), w >= 0 branch-stable (Shepperd). The reference uses an eigen-decomposition
    whose sign is arbitrary (rigid_utils.py:208-227); any consumer must treat q and -q as equal."""
    from scipy.spatial.transform import Rotation as SR

    q = SR.from_matrix(R.detach().cpu().double().reshape(-1, 3, 3).numpy()).as_quat()
    q = np.concatenate([q[:, 3:], q[:, :3]], -1).reshape(tuple(R.shape[:-2]) + (4,))
    return torch.as_tensor(q, dtype=R.dtype, device=R.device)


class Rotation:
    def __init__(self, rot_mats: torch.Tensor | None = None, quats: torch.Tensor | None = None, normalize_quats: bool = True):
        if (rot_mats is None) == (quats is None):
            raise ValueError("Exactly one input argument must be specified")
        if rot_mats is not None and rot_mats.shape[-2:] != (3, 3):
            raise ValueError("Incorrectly shaped rotation matrix or quaternion")
        if quats is not None and quats.shape[-1] != 4:
            raise ValueError("Incorrectly shaped rotation matrix or quaternion")
        if quats is not None:
            quats = quats.to(torch.float32)
            if normalize_quats:
                quats = quats / torch.linalg.norm(quats, dim=-1, keepdim=True)
        else:
            rot_mats = rot_mats.to(torch.float32)
        self._rot_mats, self._quats = rot_mats, quats

    @property
    def shape(self):
        return self._rot_mats.shape[:-2] if self._rot_mats is not None else self._quats.shape[:-1]

    @property
    def device(self):
        return (self._rot_mats if self._rot_mats is not None else self._quats).device

    def get_rot_mats(self) -> torch.Tensor:
        return self._rot_mats if self._rot_mats is not None else _quat_to_rot(self._quats)

    def get_quats(self) -> torch.Tensor:
        return self._quats if self._quats is not None else _rot_to_quat(self._rot_mats)

    def __getitem__(self, idx):
        if not isinstance(idx, tuple):
            idx = (idx,)
        if self._rot_mats is not None:
            return Rotation(rot_mats=self._rot_mats[idx + (slice(None), slice(None))])
        return Rotation(quats=self._quats[idx + (slice(None),)], normalize_quats=False)

    def to(self, device=None, dtype=None):
        f = lambda x: None if x is None else x.to(device=device)
        return Rotation(rot_mats=f(self._rot_mats), quats=f(self._quats), normalize_quats=False)


class Rigid:
    def __init__(self, rots: Rotation | None, trans: torch.Tensor | None):
        if rots is None and trans is None:
            raise ValueError("At least one argument must be specified")
        if trans is None:
            trans = torch.zeros(tuple(rots.shape) + (3,), device=rots.device)
        if rots is None:
            q = torch.zeros(tuple(trans.shape[:-1]) + (4,), device=trans.device)
            q[..., 0] = 1
            rots = Rotation(quats=q, normalize_quats=False)
        if tuple(rots.shape) != tuple(trans.shape[:-1]):
            raise ValueError("Rots and trans incompatible")
        self._rots, self._trans = rots, trans.to(torch.float32)

    @staticmethod
    def identity(shape, dtype=None, device=None, requires_grad=False, fmt="quat"):
        q = torch.zeros(tuple(shape) + (4,), device=device)
        q[..., 0] = 1
        return Rigid(Rotation(quats=q, normalize_quats=False), torch.zeros(tuple(shape) + (3,), device=device))

    @staticmethod
    def from_tensor_7(t: torch.Tensor, normalize_quats: bool = False) -> "Rigid":
        if t.shape[-1] != 7:
            raise ValueError("Incorrectly shaped input tensor")
        return Rigid(Rotation(quats=t[..., :4], normalize_quats=normalize_quats), t[..., 4:])

    def to_tensor_7(self) -> torch.Tensor:
        return torch.cat([self._rots.get_quats(), self._trans], -1)

    @property
    def shape(self):
        return self._trans.shape[:-1]

    @property
    def device(self):
        return self._trans.device

    def get_rots(self) -> Rotation:
        return self._rots

    def get_trans(self) -> torch.Tensor:
        return self._trans

    def __getitem__(self, idx):
        if not isinstance(idx, tuple):
            idx = (idx,)
        return Rigid(self._rots[idx], self._trans[idx + (slice(None),)])

    def to(self, device=None, dtype=None):
        return Rigid(self._rots.to(device=device), self._trans.to(device=device))


def rotmats_to_rigid(rotmats, trans) -> Rigid:
    return Rigid(Rotation(rot_mats=torch.as_tensor(rotmats, dtype=torch.float32)), torch.as_tensor(trans, dtype=torch.float32))
