import numpy as np


def rot_angle_between(q1, q2):
    """Angle (rad) between rotations given as (w,x,y,z) quaternions; sign-invariant, accurate for tiny angles."""
    q1 = np.asarray(q1, np.float64)
    q2 = np.asarray(q2, np.float64)
    q1 = q1 / np.linalg.norm(q1, axis=-1, keepdims=True)
    q2 = q2 / np.linalg.norm(q2, axis=-1, keepdims=True)
    d = np.minimum(np.linalg.norm(q1 - q2, axis=-1), np.linalg.norm(q1 + q2, axis=-1))
    return 4 * np.arcsin(np.clip(d / 2, 0, 1))


def bb_rmsd(a, b):
    """per-residue RMSD over N, CA, C, O (atom37 slots 0,1,2,4), unaligned (evaluation/utils/metrics.py:146-182)."""
    d = np.asarray(a, np.float64)[..., [0, 1, 2, 4], :] - np.asarray(b, np.float64)[..., [0, 1, 2, 4], :]
    return np.sqrt((d ** 2).sum(-1).mean(-1))
