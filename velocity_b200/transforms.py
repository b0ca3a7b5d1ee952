"""Drop-in for the reference's utils/transforms.py (rpy2dcm :7-23, transform :27-48, dcm2rpy :51-57,
3-element "quaternion" helpers :60-73).  Host-side scalar helpers; inside the kernels the same
Z-Y-X Euler convention is a __device__ function (velocity_b200/csrc/nls.cu, ba.cu)."""
import math

import numpy as np

from .common import norm


def _trig(rpy):
    r, p, y = float(rpy[0]), float(rpy[1]), float(rpy[2])
    return (math.sin(r), math.cos(r)), (math.sin(p), math.cos(p)), (math.sin(y), math.cos(y))


def rpy2dcm(rpy):
    """[roll, pitch, yaw] -> 3x3 direction cosine matrix (row-vector convention: X_cam = X @ C)."""
    (sr, cr), (sp, cp), (sy, cy) = _trig(rpy)
    return np.array([
        [cp * cy, sr * sp * cy - cr * sy, cr * sp * cy + sr * sy],
        [cp * sy, sr * sp * sy + cr * cy, cr * sp * sy - sr * cy],
        [-sp, sr * cp, cr * cp],
    ])


def transform(X, rpy, t):
    """X @ rpy2dcm(rpy) + t, evaluated column by column."""
    C = rpy2dcm(rpy)
    x, y, z = X[:, 0], X[:, 1], X[:, 2]
    cols = [C[0, k] * x + C[1, k] * y + C[2, k] * z + t[k] for k in range(3)]
    return np.stack(cols, axis=1)


def dcm2rpy(R):
    return np.array([math.atan(R[2, 1] / R[2, 2]), math.asin(-R[2, 0]), math.atan2(R[1, 0], R[0, 0])])


def quat2dcm(q):
    """3-element pseudo-quaternion of the reference: roll is encoded in the norm (offset 10)."""
    r = norm(q)
    return rpy2dcm([r - 10, math.asin(-q[2] / r), math.atan(q[1] / q[0])])


def dcm2quat(R):
    r, p, y = dcm2rpy(R)
    rng = r + 10
    a = rng * math.cos(p)
    return np.array([a * math.cos(y), a * math.sin(y), -rng * math.sin(p)])
