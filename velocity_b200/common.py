"""Host-side leaf helpers with the names and semantics of the reference's utils/common.py
(norm/rms/uvec :13-25, addcol0/1 :28-39, image2world :49-55, world2image :58-64, pixel2uvec
:122-126, pscale :145-147, worldPointsLicensePlate :150-156).  On the accelerated path these are
__device__ functions inside the kernels; the numpy forms here serve the thin host wrappers and the
bookkeeping the caller (vidExample.py) does between kernel calls."""
import numpy as np


def norm(x, axis=None):
    """L2 norm (note: `axis` is an axis, not a norm order -- utils/common.py:13-15)."""
    return (x * x).sum(axis) ** 0.5


def rms(x, axis=None):
    return (x * x).mean(axis) ** 0.5


def uvec(x, axis=1):
    return x / (x * x).sum(axis, keepdims=True) ** 0.5


def _addcol(x, value):
    out = np.full((x.shape[0], x.shape[1] + 1), value, x.dtype)
    out[:, :-1] = x
    return out


def addcol0(x):
    return _addcol(x, 0)


def addcol1(x):
    return _addcol(x, 1)


def pscale(p3):
    return p3[:, 0:2] / p3[:, 2:3]


def world2image(K, R, t, pw):
    cam = np.concatenate([R, t[None]]) @ K
    return pscale(addcol1(pw) @ cam)


def image2world(K, R, t, p):
    tform = np.concatenate([R[0:2, :], t[None]]) @ K
    return pscale(addcol1(p) @ np.linalg.inv(tform))


def pixel2uvec(K, p):
    q = addcol0(p - K[2, 0:2])
    q[:, 2] = K[0, 0]
    return uvec(q)


def worldPointsLicensePlate(country="EU"):
    size = [0.3725, 0.1275, 0] if country == "Chile" else [0.520, 0.110, 0]
    corners = np.array([[1, -1, 0], [1, 1, 0], [-1, 1, 0], [-1, -1, 0]], np.float32)
    return corners * np.array(size, np.float32) / 2


def cam2ned():
    return np.array([[0, 0, 1], [1, 0, 0], [0, 1, 0]])


def cc2sc(x):
    """Cartesian -> [range, elevation, azimuth] (utils/common.py:81-95); a length-3 first axis means columns."""
    x = np.asarray(x)
    cols = x.shape[0] == 3
    X, Y, Z = (x[0], x[1], x[2]) if cols else (x[:, 0], x[:, 1], x[:, 2])
    r = (X * X + Y * Y + Z * Z) ** 0.5
    parts = (r, np.arcsin(-Z / r), np.arctan2(Y, X))
    return np.stack(parts, 0 if cols else 1).astype(x.dtype)


def sc2cc(s):
    """[range, elevation, azimuth] -> Cartesian (utils/common.py:98-114)."""
    s = np.asarray(s)
    cols = s.shape[0] == 3
    r, el, az = (s[0], s[1], s[2]) if cols else (s[:, 0], s[:, 1], s[:, 2])
    a = r * np.cos(el)
    parts = (a * np.cos(az), a * np.sin(az), -r * np.sin(el))
    return np.stack(parts, 0 if cols else 1).astype(s.dtype)
