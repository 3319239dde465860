"""``grid_subsampling`` module of the reference, served by libpcrcg_b200.so on the GPU.

Mirrors cpp_wrappers.zip!cpp_wrappers/cpp_subsampling/wrapper.cpp:
  subsample_batch :62-333  ("OO|$OOfsii": points, batches positional; the rest keyword-only)
  subsample       :338-566
Inputs are coerced like PyArray_FROM_OTF(..., NPY_FLOAT/NPY_INT, NPY_IN_ARRAY): anything array-like
(including CPU torch tensors) becomes a C-contiguous float32 / int32 copy.  Results are fresh NumPy
arrays.  Failures raise RuntimeError with the reference's messages.

Deviation (documented in DESIGN.md): ``features=`` / ``classes=`` are not on the hot path
(datasets/dataloader.py:289 passes neither) and raise NotImplementedError.
"""
import ctypes as C

import numpy as np

from ..._lib import lib, check


def _as(obj, dtype, what):
    try:
        if hasattr(obj, "detach"):
            obj = obj.detach().cpu().numpy()
        return np.ascontiguousarray(np.asarray(obj), dtype=dtype)
    except Exception:
        raise RuntimeError(f"Error converting input {what} to numpy arrays of type {'float32' if dtype == np.float32 else 'int32'}")


def subsample_batch(points, batches, *, features=None, classes=None, sampleDl=0.1, method="barycenters", max_p=0, verbose=0):
    if method not in ("barycenters", "voxelcenters"):
        raise RuntimeError('Error parsing method. Valid method names are "barycenters" and "voxelcenters" ')
    if features is not None or classes is not None:
        raise NotImplementedError("pcrcg_b200: subsample_batch(features=/classes=) is outside the KPConv hot path")
    p = _as(points, np.float32, "points")
    b = _as(batches, np.int32, "batches")
    if p.ndim != 2 or p.shape[1] != 3:
        raise RuntimeError("Wrong dimensions : points.shape is not (N, 3)")
    if b.ndim > 1:
        raise RuntimeError("Wrong dimensions : batches.shape is not (B,) ")
    b = b.reshape(-1)
    L = lib()
    out = C.c_void_p()
    m = C.c_int64()
    out_lens = np.empty(len(b), np.int32)
    check(L.pcrcg_subsample_batch_host(p.ctypes.data, len(p), b.ctypes.data, len(b), float(sampleDl), int(max_p),
                                       C.byref(out), C.byref(m), out_lens.ctypes.data))
    try:
        s_points = np.ctypeslib.as_array(C.cast(out, C.POINTER(C.c_float)), shape=(m.value, 3)).copy()
    finally:
        L.pcrcg_free(out)
    return s_points, out_lens


def subsample(points, *, features=None, classes=None, sampleDl=0.1, method="barycenters", verbose=0):
    """Single-cloud variant (wrapper.cpp:338-566)."""
    p = _as(points, np.float32, "points")
    if p.ndim != 2 or p.shape[1] != 3:
        raise RuntimeError("Wrong dimensions : points.shape is not (N, 3)")
    s_points, _ = subsample_batch(p, np.array([len(p)], np.int32), features=features, classes=classes,
                                  sampleDl=sampleDl, method=method, verbose=verbose)
    return s_points
