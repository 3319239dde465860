"""``grid_subsampling`` module of the reference, served by libpcrcg_b200.so on the GPU.

Mirrors cpp_wrappers.zip!cpp_wrappers/cpp_subsampling/wrapper.cpp:
  subsample_batch :62-333  ("OO|$OOfsii": points, batches positional; the rest keyword-only)
  subsample       :338-566
Inputs are coerced like PyArray_FROM_OTF(..., NPY_FLOAT/NPY_INT, NPY_IN_ARRAY): anything array-like
(including CPU torch tensors) becomes a C-contiguous float32 / int32 copy.  Results are fresh NumPy
arrays.  Failures raise RuntimeError with the reference's messages.

``features=`` / ``classes=`` (wrapper.cpp:103-246, grid_subsampling.cpp:34-102) are served by the same library
(pcrcg_subsample_batch_ex_host): feature means and class votes per voxel, in the reference's output order, appended to the
result tuple exactly as wrapper.cpp:318-326 does.  They are not on the KPConv pyramid's path (datasets/dataloader.py:289
passes neither).  One restriction: classes with more than one column need a single cloud -- the reference slices the classes
of later clouds with a wrong end offset (grid_subsampling.cpp:157-158, out-of-bounds reads), so there is nothing to match.
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
    """-> (s_points, s_batches[, s_features][, s_classes])  (wrapper.cpp:318-326)"""
    if method not in ("barycenters", "voxelcenters"):
        raise RuntimeError('Error parsing method. Valid method names are "barycenters" and "voxelcenters" ')
    p = _as(points, np.float32, "points")
    b = _as(batches, np.int32, "batches")
    f = _as(features, np.float32, "features") if features is not None else None
    c = _as(classes, np.int32, "classes") if classes is not None else None
    if p.ndim != 2 or p.shape[1] != 3:
        raise RuntimeError("Wrong dimensions : points.shape is not (N, 3)")
    if b.ndim > 1:
        raise RuntimeError("Wrong dimensions : batches.shape is not (B,) ")
    if f is not None and f.ndim != 2:
        raise RuntimeError("Wrong dimensions : features.shape is not (N, d)")
    if c is not None and c.ndim > 2:
        raise RuntimeError("Wrong dimensions : classes.shape is not (N,) or (N, d)")
    if f is not None and f.shape[0] != len(p):
        raise RuntimeError("Wrong dimensions : features.shape is not (N, d)")
    if c is not None and (c.ndim == 0 or c.shape[0] != len(p)):
        raise RuntimeError("Wrong dimensions : classes.shape is not (N,) or (N, d)")
    b = b.reshape(-1)
    fdim = f.shape[1] if f is not None else 0
    ldim = c.shape[1] if (c is not None and c.ndim == 2) else 1
    if (f is not None and fdim < 1) or (c is not None and ldim < 1):
        raise RuntimeError("Wrong dimensions : features / classes without columns")
    if c is not None and ldim > 1 and len(b) > 1:
        raise RuntimeError("classes with more than one column are defined for a single cloud only "
                           "(the reference mis-slices them for later clouds, grid_subsampling.cpp:157-158)")
    L = lib()
    out, m = C.c_void_p(), C.c_int64()
    out_lens = np.empty(len(b), np.int32)
    if f is None and c is None:
        check(L.pcrcg_subsample_batch_host(p.ctypes.data, len(p), b.ctypes.data, len(b), float(sampleDl), int(max_p),
                                           C.byref(out), C.byref(m), out_lens.ctypes.data))
        try:
            s_points = np.ctypeslib.as_array(C.cast(out, C.POINTER(C.c_float)), shape=(m.value, 3)).copy()
        finally:
            L.pcrcg_free(out)
        return s_points, out_lens
    of, oc = C.c_void_p(), C.c_void_p()
    check(L.pcrcg_subsample_batch_ex_host(p.ctypes.data, len(p), b.ctypes.data, len(b), float(sampleDl), int(max_p),
                                          f.ctypes.data if f is not None else None, fdim, c.ctypes.data if c is not None else None, ldim,
                                          C.byref(out), C.byref(m), out_lens.ctypes.data, C.byref(of), C.byref(oc)))
    try:
        res = [np.ctypeslib.as_array(C.cast(out, C.POINTER(C.c_float)), shape=(m.value, 3)).copy(), out_lens]
        if f is not None:
            res.append(np.ctypeslib.as_array(C.cast(of, C.POINTER(C.c_float)), shape=(m.value, fdim)).copy())
        if c is not None:
            res.append(np.ctypeslib.as_array(C.cast(oc, C.POINTER(C.c_int32)), shape=(m.value, ldim)).copy())
    finally:
        for ptr in (out, of, oc):
            if ptr.value:
                L.pcrcg_free(ptr)
    return tuple(res)


def subsample(points, *, features=None, classes=None, sampleDl=0.1, method="barycenters", verbose=0):
    """Single-cloud variant (wrapper.cpp:338-566): -> s_points, or (s_points[, s_features][, s_classes])  (:546-553)."""
    p = _as(points, np.float32, "points")
    if p.ndim != 2 or p.shape[1] != 3:
        raise RuntimeError("Wrong dimensions : points.shape is not (N, 3)")
    res = subsample_batch(p, np.array([len(p)], np.int32), features=features, classes=classes, sampleDl=sampleDl, method=method,
                          verbose=verbose)
    if len(res) == 2:
        return res[0]
    return (res[0],) + tuple(res[2:])
