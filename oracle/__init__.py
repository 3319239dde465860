"""oracle -- TEST INFRASTRUCTURE ONLY.

CPU checkers for the CUDA hot path:

* ``port``  : plain-C restatement (oracle/port.c -> liboracle_port.so) + PyTorch/NumPy restatements
              (oracle/blocks_port.py, oracle/projection_port.py).  Travels to the GPU box.
* ``ref``   : the UNMODIFIED reference C++ core compiled behind oracle/ref_shim.cpp
              (oracle/_ref/libpcrcg_ref.so).  Built only where /root/reference exists; the built
              library travels to the GPU box.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this package.  Nothing under ``pcrcg_b200/`` does: the product path fails loudly
when its CUDA library is missing instead of falling back to anything here.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PORT_SO = os.path.join(_HERE, "liboracle_port.so")
_REF_SO = os.path.join(_HERE, "_ref", "libpcrcg_ref.so")

_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")


def build(ref=True):
    """(Re)build the checkers with oracle/Makefile.  Building the checker is not using it."""
    subprocess.check_call(["make", "-s", "-C", _HERE, "port"])
    if ref:
        subprocess.check_call(["make", "-s", "-C", _HERE, "ref"])


def _f32(a, cols=None):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float32))
    if cols is not None:
        assert a.ndim == 2 and a.shape[1] == cols, a.shape
    return a


def _i32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.int32)).reshape(-1)


def _subsample_ex(fn, has_cap, points, batches, features, classes, sampleDl, max_p):
    p, b = _f32(points, 3), _i32(batches)
    n1 = max(len(p), 1)
    f = None if features is None else _f32(features)
    c = None if classes is None else np.ascontiguousarray(np.asarray(classes, dtype=np.int32))
    fdim = 0 if f is None else f.shape[1]
    if c is not None and c.ndim == 1:
        c = c.reshape(-1, 1)
    ldim = 1 if c is None else c.shape[1]
    out, ol = np.empty((n1, 3), np.float32), np.empty(len(b), np.int32)
    of = np.empty((n1, max(fdim, 1)), np.float32)
    oc = np.empty((n1, ldim), np.int32)
    args = [p, len(p), b, len(b), sampleDl, max_p, None if f is None else f.ctypes.data, fdim, None if c is None else c.ctypes.data, ldim, out]
    args += ([len(out)] if has_cap else []) + [ol, of.ctypes.data, oc.ctypes.data]
    m = fn(*args)
    if m == -2:
        raise ValueError("classes with more than one column and more than one cloud: undefined in the reference (grid_subsampling.cpp:157-158)")
    assert m >= 0
    res = [out[:m].copy(), ol]
    if f is not None:
        res.append(of[:m, :fdim].copy())
    if c is not None:
        res.append(oc[:m].copy())
    return tuple(res)


# ------------------------------------------------------------------------------------------------
class _Port:
    def __init__(self):
        if not os.path.exists(_PORT_SO):
            build(ref=False)
        L = C.CDLL(_PORT_SO)
        L.oracle_subsample_batch.restype = C.c_int64
        L.oracle_subsample_batch.argtypes = [_f32p, C.c_int64, _i32p, C.c_int32, C.c_float, C.c_int32, _f32p, _i32p]
        L.oracle_subsample_batch_ex.restype = C.c_int64
        L.oracle_subsample_batch_ex.argtypes = [_f32p, C.c_int64, _i32p, C.c_int32, C.c_float, C.c_int32, C.c_void_p, C.c_int32,
                                                C.c_void_p, C.c_int32, _f32p, _i32p, C.c_void_p, C.c_void_p]
        L.oracle_label_vote.restype = C.c_int32
        L.oracle_label_vote.argtypes = [_i32p, C.c_int64]
        L.oracle_voxel_keys.restype = None
        L.oracle_voxel_keys.argtypes = [_f32p, C.c_int64, C.c_float, _u64p, _f32p, _u64p]
        L.oracle_radius_count.restype = C.c_int32
        L.oracle_radius_count.argtypes = [_f32p, C.c_int64, _f32p, C.c_int64, _i32p, _i32p, C.c_int32, C.c_float, _i32p]
        L.oracle_radius_fill.restype = None
        L.oracle_radius_fill.argtypes = [_f32p, C.c_int64, _f32p, C.c_int64, _i32p, _i32p, C.c_int32, C.c_float, C.c_int32, _i32p]
        L.oracle_canonicalise_rows.restype = C.c_int64
        L.oracle_canonicalise_rows.argtypes = [_f32p, C.c_int64, _f32p, C.c_int64, C.c_int32, _i32p]
        L.oracle_bucket_schedule.restype = C.c_int
        L.oracle_bucket_schedule.argtypes = [_u64p, C.c_int]
        self.L = L

    # grid_subsampling.cpp:109-211
    def subsample_batch(self, points, batches, sampleDl=0.1, max_p=0):
        p, b = _f32(points, 3), _i32(batches)
        out = np.empty((max(len(p), 1), 3), np.float32)
        ol = np.empty(len(b), np.int32)
        m = self.L.oracle_subsample_batch(p, len(p), b, len(b), sampleDl, max_p, out, ol)
        return out[:m].copy(), ol

    # grid_subsampling.cpp:109-211 with features and / or classes -> (points, lens[, features][, classes]) like wrapper.cpp:318-326
    def subsample_batch_ex(self, points, batches, features=None, classes=None, sampleDl=0.1, max_p=0):
        return _subsample_ex(self.L.oracle_subsample_batch_ex, False, points, batches, features, classes, sampleDl, max_p)

    def label_vote(self, labels):
        l = _i32(labels)
        return int(self.L.oracle_label_vote(l, len(l)))

    def voxel_keys(self, points, sampleDl):
        p = _f32(points, 3)
        keys = np.empty(len(p), np.uint64)
        org = np.empty(3, np.float32)
        nxny = np.empty(2, np.uint64)
        self.L.oracle_voxel_keys(p, len(p), sampleDl, keys, org, nxny)
        return keys, org, nxny

    def radius_counts(self, queries, supports, q_batches, s_batches, radius):
        q, s, ql, sl = _f32(queries, 3), _f32(supports, 3), _i32(q_batches), _i32(s_batches)
        counts = np.empty(len(q), np.int32)
        mx = self.L.oracle_radius_count(q, len(q), s, len(s), ql, sl, len(ql), radius, counts)
        return counts, int(mx)

    # neighbors.cpp:211-332 with the canonical (d2, index) order; limit<=0 -> full width max_count
    def batch_query(self, queries, supports, q_batches, s_batches, radius, limit=0):
        q, s, ql, sl = _f32(queries, 3), _f32(supports, 3), _i32(q_batches), _i32(s_batches)
        counts, mx = self.radius_counts(q, s, ql, sl, radius)
        width = mx if limit <= 0 else min(limit, mx)
        out = np.empty((len(q), width), np.int32)
        self.L.oracle_radius_fill(q, len(q), s, len(s), ql, sl, len(ql), radius, width, out)
        return out

    def canonicalise_rows(self, queries, supports, rows):
        q, s = _f32(queries, 3), _f32(supports, 3)
        rows = np.ascontiguousarray(rows, dtype=np.int32).copy()
        changed = self.L.oracle_canonicalise_rows(q, len(q), s, len(s), rows.shape[1], rows)
        return rows, int(changed)

    def bucket_schedule(self):
        out = np.empty(64, np.uint64)
        n = self.L.oracle_bucket_schedule(out, 64)
        return out[:n].copy()


class _Ref:
    """The real reference core (kd-tree search, unordered_map subsampling)."""

    def __init__(self):
        L = C.CDLL(_REF_SO)
        L.ref_subsample_batch.restype = C.c_long
        L.ref_subsample_batch.argtypes = [_f32p, C.c_long, _i32p, C.c_int, C.c_float, C.c_int, _f32p, C.c_long, _i32p]
        L.ref_subsample_batch_ex.restype = C.c_long
        L.ref_subsample_batch_ex.argtypes = [_f32p, C.c_long, _i32p, C.c_int, C.c_float, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                             _f32p, C.c_long, _i32p, C.c_void_p, C.c_void_p]
        L.ref_label_vote.restype = C.c_int
        L.ref_label_vote.argtypes = [_i32p, C.c_long]
        L.ref_batch_query.restype = C.c_long
        L.ref_batch_query.argtypes = [_f32p, C.c_long, _f32p, C.c_long, _i32p, _i32p, C.c_int, C.c_float]
        L.ref_batch_query_fetch.restype = None
        L.ref_batch_query_fetch.argtypes = [_i32p]
        L.ref_bucket_schedule.restype = C.c_int
        L.ref_bucket_schedule.argtypes = [C.c_long, _i64p, C.c_int]
        self.L = L

    def subsample_batch(self, points, batches, sampleDl=0.1, max_p=0):
        p, b = _f32(points, 3), _i32(batches)
        out = np.empty((max(len(p), 1), 3), np.float32)
        ol = np.empty(len(b), np.int32)
        m = self.L.ref_subsample_batch(p, len(p), b, len(b), sampleDl, max_p, out, len(out), ol)
        assert m >= 0
        return out[:m].copy(), ol

    def subsample_batch_ex(self, points, batches, features=None, classes=None, sampleDl=0.1, max_p=0):
        return _subsample_ex(self.L.ref_subsample_batch_ex, True, points, batches, features, classes, sampleDl, max_p)

    def label_vote(self, labels):
        l = _i32(labels)
        return int(self.L.ref_label_vote(l, len(l)))

    def batch_query(self, queries, supports, q_batches, s_batches, radius):
        """Raw reference output [Nq, max_count] (tie order = kd-tree traversal + unstable sort)."""
        q, s, ql, sl = _f32(queries, 3), _f32(supports, 3), _i32(q_batches), _i32(s_batches)
        w = self.L.ref_batch_query(q, len(q), s, len(s), ql, sl, len(ql), radius)
        out = np.empty((len(q), w), np.int32)
        self.L.ref_batch_query_fetch(out)
        return out

    def bucket_schedule(self, n):
        out = np.empty(64, np.int64)
        k = self.L.ref_bucket_schedule(n, out, 64)
        return out[:k].copy()


_port = None
_ref = None


def port():
    global _port
    if _port is None:
        _port = _Port()
    return _port


def have_ref():
    return os.path.exists(_REF_SO)


def ref():
    global _ref
    if _ref is None:
        if not have_ref():
            raise RuntimeError("oracle/_ref/libpcrcg_ref.so not built (needs /root/reference)")
        _ref = _Ref()
    return _ref


def ref_batch_query_canonical(queries, supports, q_batches, s_batches, radius, limit=0):
    """Reference search -> canonical (d2, index) tie order -> python-side truncation
    (datasets/dataloader.py:66-69).  Returns (rows, n_rows_reordered)."""
    raw = ref().batch_query(queries, supports, q_batches, s_batches, radius)
    rows, changed = port().canonicalise_rows(queries, supports, raw)
    if limit > 0:
        rows = np.ascontiguousarray(rows[:, :limit])
    return rows, changed
