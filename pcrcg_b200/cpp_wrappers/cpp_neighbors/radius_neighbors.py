"""``radius_neighbors`` module of the reference, served by libpcrcg_b200.so on the GPU.

Mirrors cpp_wrappers/cpp_neighbors/wrapper.cpp:58-238 (batch_query, format "OOOO|$f").  Returns a
fresh int32 NumPy array [Nq, max_count]; neighbours in ascending (d2, index); shadow index = Ns.
"""
import ctypes as C

import numpy as np

from ..._lib import lib, check


def _as(obj, dtype, msg):
    try:
        if hasattr(obj, "detach"):
            obj = obj.detach().cpu().numpy()
        return np.ascontiguousarray(np.asarray(obj), dtype=dtype)
    except Exception:
        raise RuntimeError(msg)


def batch_query(queries, supports, q_batches, s_batches, *, radius=0.1):
    q = _as(queries, np.float32, "Error converting query points to numpy arrays of type float32")
    s = _as(supports, np.float32, "Error converting support points to numpy arrays of type float32")
    ql = _as(q_batches, np.int32, "Error converting query batches to numpy arrays of type int32")
    sl = _as(s_batches, np.int32, "Error converting support batches to numpy arrays of type int32")
    if q.ndim != 2 or q.shape[1] != 3:
        raise RuntimeError("Wrong dimensions : query.shape is not (N, 3)")
    if s.ndim != 2 or s.shape[1] != 3:
        raise RuntimeError("Wrong dimensions : support.shape is not (N, 3)")
    if ql.ndim > 1:
        raise RuntimeError("Wrong dimensions : queries_batches.shape is not (B,) ")
    if sl.ndim > 1:
        raise RuntimeError("Wrong dimensions : supports_batches.shape is not (B,) ")
    ql, sl = ql.reshape(-1), sl.reshape(-1)
    if len(ql) != len(sl):
        raise RuntimeError("Wrong number of batch elements: different for queries and supports ")
    L = lib()
    out = C.c_void_p()
    w = C.c_int32()
    check(L.pcrcg_batch_query_host(q.ctypes.data, len(q), s.ctypes.data, len(s), ql.ctypes.data, sl.ctypes.data, len(ql),
                                   float(radius), 0, C.byref(out), C.byref(w)))
    try:
        rows = np.ctypeslib.as_array(C.cast(out, C.POINTER(C.c_int32)), shape=(len(q), w.value)).copy()
    finally:
        L.pcrcg_free(out)
    return rows
