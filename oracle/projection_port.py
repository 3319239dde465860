"""oracle/projection_port.py -- TEST INFRASTRUCTURE ONLY.

Restatement of the reference colour path: ``projection.py:31-61`` (index generation, in C:
oracle/port.c::oracle_projection) and the gather/scatter of 2D features into the 129-wide input
rows (``models/architectures.py:273-307,360-370``, img_num == 2 branch), in NumPy.
"""
import ctypes as C

import numpy as np

from . import port, _f32p

_i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")


def projection(points, depth_map, world2camera, intrinsics, thresh=0.1):
    """-> (inds2d int64 [M,2] (x,y), inds3d int64 [M])."""
    L = port().L
    L.oracle_projection.restype = C.c_int64
    L.oracle_projection.argtypes = [_f32p, C.c_int64, _f32p, C.c_int32, C.c_int32, _f32p, _f32p,
                                    C.c_float, _i64p, _i64p]
    p = np.ascontiguousarray(points, np.float32)
    d = np.ascontiguousarray(depth_map, np.float32)
    d = d.reshape(d.shape[-2], d.shape[-1])
    w = np.ascontiguousarray(world2camera, np.float32).reshape(16)
    k = np.ascontiguousarray(intrinsics, np.float32)
    if k.shape == (3, 3):                       # projection.py:22-25
        k4 = np.eye(4, dtype=np.float32)
        k4[:3, :3] = k
        k = k4
    k = k.reshape(16)
    i2 = np.empty((max(len(p), 1), 2), np.int64)
    i3 = np.empty(max(len(p), 1), np.int64)
    m = L.oracle_projection(p, len(p), d, d.shape[0], d.shape[1], w, k, np.float32(thresh), i2, i3)
    return i2[:m].copy(), i3[:m].copy()


def scatter_image_features(n_points, views):
    """models/architectures.py:273-307,360-370.

    views: list, in WRITE order (the reference writes image 2 first, image 1 last, per cloud), of
    (feature2d [C,H,W], valid_map [H,W] or None, inds2d [M,2], inds3d [M] global row indices).
    Returns x [n_points, C+1]: ones everywhere, then rows overwritten by [feat*valid, 1]."""
    c = views[0][0].shape[0]
    x = np.ones((n_points, c + 1), np.float32)          # features ones [N,1] .repeat(1,129)
    for f2d, valid, i2, i3 in views:
        f = f2d if valid is None else f2d * valid[None]
        rows = f[:, i2[:, 1], i2[:, 0]].T               # [M,C]
        x[i3, :c] = rows
        x[i3, c] = 1.0
    return x
