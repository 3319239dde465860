"""Device-level operators: torch CUDA tensors in, torch CUDA tensors out, computed by
libpcrcg_b200.so on the tensor's device and torch's current stream.  PyTorch is used for device
memory and streams only."""
import torch

from ._lib import lib, check


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("pcrcg_b200.ops: CUDA tensors required (there is no CPU path)")


def _f32c(t):
    return t.contiguous() if t.dtype == torch.float32 else t.float().contiguous()


def _i32c(t):
    return t.contiguous() if t.dtype == torch.int32 else t.to(torch.int32).contiguous()


def _ws(nbytes, device):
    return torch.empty(int(nbytes), dtype=torch.uint8, device=device)


# Derived data that rides on a tensor as Python attributes (the bf16 planes of its values, its InstanceNorm statistics, its
# row-positive flags) is stamped with the tensor's version counter and ignored once an in-place update has changed the
# values.  Ops of this module that write in place through the C library (which the counter does not see) strip it.
_DERIVED = ("_pcrcg_stats", "_pcrcg_split", "_pcrcg_rowpos", "_pcrcg_wsplit")


def _attach(t, name, value):
    setattr(t, name, (value, t._version))


def _attached(t, name):
    if isinstance(t, PlaneTensor):
        return getattr(t, name, None)
    v = getattr(t, name, None)
    if v is None:
        return None
    value, version = v
    return value if version == t._version else None


attached = _attached          # public: derived data of a tensor if still valid (tests, diagnostics)


def _strip_derived(t):
    for name in _DERIVED:
        if hasattr(t, name):
            delattr(t, name)


def subsample_batch(points, lens, sampleDl, max_p=0):
    """points [N,3] f32 cuda, lens [B] i32 cuda -> (s_points [M,3], s_lens [B] i32 cuda).
    One host sync (to learn M)."""
    _need_cuda(points, lens)
    points, lens = _f32c(points), _i32c(lens)
    n, nb = points.shape[0], lens.shape[0]
    L = lib()
    with torch.cuda.device(points.device):
        ws = _ws(L.pcrcg_subsample_ws_bytes(n, nb), points.device)
        out = torch.empty((max(n, 1), 3), dtype=torch.float32, device=points.device)
        out_lens = torch.empty(nb, dtype=torch.int32, device=points.device)
        check(L.pcrcg_subsample_batch_dev(points.data_ptr(), n, lens.data_ptr(), nb, float(sampleDl), int(max_p),
                                          out.data_ptr(), out_lens.data_ptr(), ws.data_ptr(), ws.numel(), _stream()))
        tot = torch.empty(1, dtype=torch.int32, device=points.device)
        gs = torch.empty(nb + 1, dtype=torch.int32, device=points.device)
        check(L.pcrcg_group_starts_dev(out_lens.data_ptr(), nb, 1, gs.data_ptr(), tot.data_ptr(), _stream()))
        m = int(tot.item())
    return out[:m], out_lens


def subsample_batch_ex(points, lens, sampleDl, max_p=0, features=None, classes=None):
    """Grid subsampling with per-point features [N, d] f32 and / or classes [N] or [N, d] i32 (grid_subsampling.cpp:34-102), all
    cuda -> (s_points, s_lens[, s_features [M, d]][, s_classes [M, d]]) in the reference's tuple order (wrapper.cpp:318-326).
    One host sync (M and the status word of the class votes)."""
    if features is None and classes is None:
        return subsample_batch(points, lens, sampleDl, max_p)
    _need_cuda(points, lens)
    points, lens = _f32c(points), _i32c(lens)
    n, nb = points.shape[0], lens.shape[0]
    f = c = None
    fdim, ldim = 0, 1
    if features is not None:
        _need_cuda(features)
        f = _f32c(features)
        if f.dim() != 2 or f.shape[0] != n or f.shape[1] < 1:
            raise RuntimeError("Wrong dimensions : features.shape is not (N, d)")
        fdim = f.shape[1]
    if classes is not None:
        _need_cuda(classes)
        c = _i32c(classes)
        if c.dim() not in (1, 2) or c.shape[0] != n or (c.dim() == 2 and c.shape[1] < 1):
            raise RuntimeError("Wrong dimensions : classes.shape is not (N,) or (N, d)")
        ldim = c.shape[1] if c.dim() == 2 else 1
    L = lib()
    dev = points.device
    with torch.cuda.device(dev):
        ws = _ws(L.pcrcg_subsample_ex_ws_bytes(n, nb, fdim, ldim if c is not None else 0), dev)
        out = torch.empty((max(n, 1), 3), dtype=torch.float32, device=dev)
        out_lens = torch.empty(nb, dtype=torch.int32, device=dev)
        of = torch.empty((max(n, 1), fdim), dtype=torch.float32, device=dev) if f is not None else None
        oc = torch.empty((max(n, 1), ldim), dtype=torch.int32, device=dev) if c is not None else None
        tot = torch.zeros(2, dtype=torch.int32, device=dev)           # [M, status]
        check(L.pcrcg_subsample_batch_ex_dev(points.data_ptr(), n, lens.data_ptr(), nb, float(sampleDl), int(max_p),
                                             f.data_ptr() if f is not None else None, fdim, c.data_ptr() if c is not None else None, ldim,
                                             out.data_ptr(), out_lens.data_ptr(), of.data_ptr() if of is not None else None,
                                             oc.data_ptr() if oc is not None else None, tot[1:].data_ptr(), ws.data_ptr(), ws.numel(),
                                             _stream()))
        gs = torch.empty(nb + 1, dtype=torch.int32, device=dev)
        check(L.pcrcg_group_starts_dev(out_lens.data_ptr(), nb, 1, gs.data_ptr(), tot.data_ptr(), _stream()))
        m, status = tot.tolist()
    if status != 0:
        raise RuntimeError("subsample_batch: a voxel holds more than 64 distinct labels in one class column (the tie order of the "
                           "reference's unordered_map is modelled up to 64)")
    res = [out[:m], out_lens]
    if of is not None:
        res.append(of[:m])
    if oc is not None:
        res.append(oc[:m])
    return tuple(res)


class RadiusGrid:
    """Support cloud binned for radius search (pcrcg_radius_build_dev); query it any number of times."""

    def __init__(self, supports, s_lens, radius):
        _need_cuda(supports, s_lens)
        self.supports, self.s_lens = _f32c(supports), _i32c(s_lens)
        self.radius = float(radius)
        self.ns, self.nb = self.supports.shape[0], self.s_lens.shape[0]
        L = lib()
        with torch.cuda.device(self.supports.device):
            self.ws = _ws(L.pcrcg_radius_ws_bytes(self.ns, self.ns, self.nb), self.supports.device)
            check(L.pcrcg_radius_build_dev(self.supports.data_ptr(), self.ns, self.s_lens.data_ptr(), self.nb, self.radius,
                                           self.ws.data_ptr(), self.ws.numel(), _stream()))

    def query(self, queries, q_lens, width, want_counts=True):
        """-> rows [Nq,width] i32 (None if width == 0), counts [Nq] i32, max_count [1] i32 (device).
        Cell-centric search (pcrcg_radius_query_cells_dev); ``cell_centric(False)`` selects the one-warp-per-query kernel."""
        _need_cuda(queries, q_lens)
        queries, q_lens = _f32c(queries), _i32c(q_lens)
        nq = queries.shape[0]
        dev = queries.device
        L = lib()
        with torch.cuda.device(dev):
            rows = torch.empty((nq, width), dtype=torch.int32, device=dev) if width > 0 else None
            counts = torch.empty(nq, dtype=torch.int32, device=dev) if want_counts else None
            mx = torch.empty(1, dtype=torch.int32, device=dev)            # cleared by the query call
            rp = rows.data_ptr() if rows is not None else None
            cp = counts.data_ptr() if counts is not None else None
            if _cell_centric:
                same = queries.data_ptr() == self.supports.data_ptr() and nq == self.ns
                qws = None if same else _ws(L.pcrcg_radius_query_ws_bytes(nq, self.nb), dev)
                check(L.pcrcg_radius_query_cells_dev(queries.data_ptr(), nq, q_lens.data_ptr(), self.ns, self.nb, self.radius,
                                                     int(width), int(width), rp, cp, mx.data_ptr(), self.ws.data_ptr(), self.ws.numel(),
                                                     qws.data_ptr() if qws is not None else None, qws.numel() if qws is not None else 0,
                                                     1 if same else 0, _stream()))
            else:
                check(L.pcrcg_radius_query_dev(queries.data_ptr(), nq, q_lens.data_ptr(), self.ns, self.nb, self.radius,
                                               int(width), int(width), rp, cp, mx.data_ptr(), self.ws.data_ptr(), self.ws.numel(), _stream()))
        return rows, counts, mx


_cell_centric = True


def cell_centric(on):
    """A/B switch of the radius search kernel (True: one warp per occupied query cell, False: one warp per query)."""
    global _cell_centric
    _cell_centric = bool(on)


def batch_query(queries, supports, q_lens, s_lens, radius, limit=0):
    """Reference semantics on device tensors: rows [Nq, min(limit, max_count)] (limit<=0: max_count).
    One host sync (to learn max_count)."""
    g = RadiusGrid(supports, s_lens, radius)
    if limit > 0:
        rows, _, mx = g.query(queries, q_lens, limit, want_counts=False)
        w = min(int(mx.item()), limit)
        return rows if w == limit else rows[:, :w].contiguous()
    _, _, mx = g.query(queries, q_lens, 0, want_counts=False)
    w = int(mx.item())
    rows, _, _ = g.query(queries, q_lens, w, want_counts=False)
    return rows


# =================================================================================================
# KPConv operator library (models/blocks.py)
# =================================================================================================
def _idx(t):
    """index tensors: int32 (native) or int64 (the reference's .long() lists) are both taken as is."""
    if t.dtype not in (torch.int32, torch.int64):
        t = t.to(torch.int64)
    if t.dim() != 2 or t.stride(1) != 1:
        t = t.contiguous()
    return t, int(t.dtype == torch.int64), t.shape[1], t.stride(0)


_fuse_stats = True


def fuse_statistics(on):
    """True (default): the tensor-core contractions accumulate the InstanceNorm statistics of their output in
    the epilogue; False: separate statistics pass (pcrcg_colstats_dev) -- A/B switch for measurements."""
    global _fuse_stats
    _fuse_stats = bool(on)


def _stats_begin(n, cout, stat_segments, device):
    """-> (seg_starts, acc) for an epilogue statistics sink, or (None, None)"""
    if stat_segments is False or not _fuse_stats or _force_simt or cout % 16 != 0 or n < 1:
        return None, None
    seg = _seg_starts(n, None if stat_segments is True else stat_segments, device)
    return seg, torch.empty((seg.shape[0] - 1, 2, cout), dtype=torch.float64, device=device)      # cleared by the *_stats_dev call


def _stats_end(out, seg, acc, eps=1e-5):
    nseg, _, c = acc.shape
    mean = torch.empty((nseg, c), dtype=torch.float32, device=out.device)
    rstd = torch.empty((nseg, c), dtype=torch.float32, device=out.device)
    check(lib().pcrcg_colstats_final_dev(acc.data_ptr(), seg.data_ptr(), nseg, c, float(eps), mean.data_ptr(), rstd.data_ptr(), _stream()))
    _attach(out, "_pcrcg_stats", (mean, rstd, seg, float(eps)))


def kpconv_forward(q_pts, s_pts, neighb_inds, x, kernel_points, weights, KP_extent, stat_segments=False):
    """models/blocks.py:229-374 (rigid, linear, sum).  -> [Nq, Cout] float32.
    stat_segments: False = plain; True / int32 row starts = also accumulate the InstanceNorm statistics of the
    result in the contraction epilogue (attached as ``out._pcrcg_stats`` for :func:`instance_norm_act`)."""
    _need_cuda(q_pts, s_pts, neighb_inds, x, kernel_points, weights)
    planes = isinstance(x, PlaneTensor)
    if planes and (_force_simt or x._pcrcg_rowpos is None or weights.shape[2] % 16 != 0 or stat_segments is False or not _fuse_stats):
        x, planes = x.dense(), False            # paths that read the fp32 rows
    q_pts, s_pts = _f32c(q_pts), _f32c(s_pts)
    if not planes:
        x = _f32c(x)
    kernel_points, weights = _f32c(kernel_points), _f32c(weights)
    idx, is64, H, stride = _idx(neighb_inds)
    nq, ns, cin = q_pts.shape[0], s_pts.shape[0], x.shape[1]
    K, cin_w, cout = weights.shape
    if cin_w != cin or x.shape[0] != ns or idx.shape[0] != nq or kernel_points.shape[0] != K:
        raise RuntimeError("kpconv_forward: inconsistent shapes")
    L = lib()
    dev = x.device
    with torch.cuda.device(dev):
        out = torch.empty((nq, cout), dtype=torch.float32, device=dev)
        ws = _ws(L.pcrcg_kpconv_ws_bytes(nq, ns, cin, K), dev)
        sp = _attached(x, "_pcrcg_split")
        xptr = None if planes else x.data_ptr()
        seg, acc = _stats_begin(nq, cout, stat_segments, dev)
        if planes and (acc is None or K * cin < 16):
            raise RuntimeError("kpconv_forward: planes-only features need the tensor-core path")
        if acc is not None and K * cin >= 16:
            hi, lo, ld = sp if sp is not None else (None, None, 0)
            rp = _attached(x, "_pcrcg_rowpos") if sp is not None else None
            check(L.pcrcg_kpconv_forward_stats_dev(q_pts.data_ptr(), nq, s_pts.data_ptr(), ns, idx.data_ptr(), is64, H, stride,
                                                   xptr, hi.data_ptr() if hi is not None else None,
                                                   lo.data_ptr() if lo is not None else None, ld,
                                                   rp.data_ptr() if rp is not None else None, cin, kernel_points.data_ptr(), K,
                                                   float(KP_extent), weights.data_ptr(), cout, out.data_ptr(), ws.data_ptr(),
                                                   ws.numel(), seg.data_ptr(), seg.shape[0] - 1, acc.data_ptr(), _stream()))
            _stats_end(out, seg, acc)
        elif sp is not None and not _force_simt:
            hi, lo, ld = sp
            rp = _attached(x, "_pcrcg_rowpos")
            check(L.pcrcg_kpconv_forward_split_dev(q_pts.data_ptr(), nq, s_pts.data_ptr(), ns, idx.data_ptr(), is64, H, stride,
                                                   x.data_ptr(), hi.data_ptr(), lo.data_ptr(), ld,
                                                   rp.data_ptr() if rp is not None else None, cin, kernel_points.data_ptr(), K,
                                                   float(KP_extent), weights.data_ptr(), cout, out.data_ptr(), ws.data_ptr(),
                                                   ws.numel(), _stream()))
        else:
            check(L.pcrcg_kpconv_forward_dev(q_pts.data_ptr(), nq, s_pts.data_ptr(), ns, idx.data_ptr(), is64, H, stride,
                                             x.data_ptr(), cin, kernel_points.data_ptr(), K, float(KP_extent), weights.data_ptr(),
                                             cout, out.data_ptr(), ws.data_ptr(), ws.numel(), _stream()))
    return out


_force_simt = False


def _split_planes(n, c, device):
    ld = (c + 7) // 8 * 8
    return (torch.empty((n, ld), dtype=torch.bfloat16, device=device), torch.empty((n, ld), dtype=torch.bfloat16, device=device), ld)


def linear(x, weight, stat_segments=False):
    """nn.Linear(bias=False): x [N,Cin] @ weight[Cout,Cin]^T   (models/blocks.py:490,497).
    If the producer of ``x`` attached its bf16 (hi, lo) planes (``x._pcrcg_split``) the tensor-core
    contraction consumes them directly.  stat_segments: see :func:`kpconv_forward`."""
    _need_cuda(x, weight)
    weight = _f32c(weight)
    cout = weight.shape[0]
    if isinstance(x, PlaneTensor):
        if _force_simt or cout % 16 != 0 or x.shape[1] < 16:
            x = x.dense()
    else:
        x = _f32c(x)
    n, cin = x.shape
    out = torch.empty((n, cout), dtype=torch.float32, device=x.device)
    L = lib()
    sp = _attached(x, "_pcrcg_split")
    with torch.cuda.device(x.device):
        if sp is not None and not _force_simt and cout % 16 == 0 and cin >= 16 and n >= 1:
            hi, lo, ld = sp
            # the weights' bf16 planes are cached on the weight tensor (version-stamped: an in-place update re-splits them)
            wsp = _attached(weight, "_pcrcg_wsplit")
            if wsp is None or wsp[2] != ld or wsp[0].device != x.device:
                bh, bl, _ = _split_planes(cout, cin, x.device)
                if bh.shape[1] != ld:
                    bh = torch.empty((cout, ld), dtype=torch.bfloat16, device=x.device)
                    bl = torch.empty((cout, ld), dtype=torch.bfloat16, device=x.device)
                check(L.pcrcg_split_bf16_dev(weight.data_ptr(), cin, cout, cin, bh.data_ptr(), bl.data_ptr(), ld, _stream()))
                _attach(weight, "_pcrcg_wsplit", (bh, bl, ld))
            else:
                bh, bl, _ = wsp
            seg, acc = _stats_begin(n, cout, stat_segments, x.device)
            if acc is not None:
                check(L.pcrcg_gemm_bf16x3_stats_dev(hi.data_ptr(), lo.data_ptr(), bh.data_ptr(), bl.data_ptr(), ld, out.data_ptr(), cout, n, cout,
                                                    cin, None, seg.data_ptr(), seg.shape[0] - 1, acc.data_ptr(), _stream()))
                _stats_end(out, seg, acc)
            else:
                check(L.pcrcg_gemm_bf16x3_dev(hi.data_ptr(), lo.data_ptr(), bh.data_ptr(), bl.data_ptr(), ld, out.data_ptr(), cout, n, cout, cin,
                                              None, _stream()))
        else:
            check(L.pcrcg_gemm_dev(x.data_ptr(), cin, weight.data_ptr(), cin, 1, out.data_ptr(), cout, n, cout, cin, None, _stream()))
    return out


def matmul(a, b):
    """a [M,K] @ b [K,N] (test helper for the contraction kernels)."""
    a, b = _f32c(a), _f32c(b)
    m, k = a.shape
    n = b.shape[1]
    out = torch.empty((m, n), dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device):
        check(lib().pcrcg_gemm_dev(a.data_ptr(), k, b.data_ptr(), n, 0, out.data_ptr(), n, m, n, k, None, _stream()))
    return out


def _seg_starts(n, segments, device):
    """segments: None (one group = all rows, the reference) or an int32 device tensor [nseg+1] of row starts."""
    if segments is None:
        return torch.tensor([0, n], dtype=torch.int32, device=device)
    return _i32c(segments)


def column_stats(x, segments=None, eps=1e-5):
    x = _f32c(x)
    n, c = x.shape
    seg = _seg_starts(n, segments, x.device)
    nseg = seg.shape[0] - 1
    mean = torch.empty((nseg, c), dtype=torch.float32, device=x.device)
    rstd = torch.empty((nseg, c), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        check(lib().pcrcg_colstats_dev(x.data_ptr(), n, c, seg.data_ptr(), nseg, float(eps), mean.data_ptr(), rstd.data_ptr(), _stream()))
    return mean, rstd, seg


class PlaneTensor:
    """Features that exist ONLY as bf16 (hi, lo) planes (x = hi + lo), produced by ``instance_norm_act(planes_only=True)``
    for results consumed solely by tensor-core contractions (a ResnetBottleneck block's unary1 and KPConv-norm outputs):
    the fp32 copy is never written.  Accepted by :func:`linear` and :func:`kpconv_forward`; ``dense()`` rebuilds fp32."""

    def __init__(self, hi, lo, ld, n, c, rowpos=None):
        self._pcrcg_split = (hi, lo, ld)
        self._pcrcg_rowpos = rowpos
        self.shape = (n, c)
        self.device = hi.device
        self.is_cuda = True
        self.dtype = torch.float32

    def dense(self):
        hi, lo, _ = self._pcrcg_split
        return (hi[:, :self.shape[1]].float() + lo[:, :self.shape[1]].float()).contiguous()


def _stats_of(x, segments, eps):
    """statistics attached by the producing contraction (same eps, same segment tensor) or a pass over x"""
    st = _attached(x, "_pcrcg_stats")
    if st is not None and st[3] == float(eps) and x.dtype == torch.float32 and x.is_contiguous():
        mean, rstd, seg, _ = st
        if (segments is None and seg.shape[0] == 2) or (segments is not None and seg.data_ptr() == _i32c(segments).data_ptr()):
            return mean, rstd, seg
    return column_stats(x, segments, eps)


def instance_norm_act(x, segments=None, slope=None, shortcut=None, shortcut_norm=False, eps=1e-5, emit_split=False,
                      emit_rowpos=False, planes_only=False):
    """act(IN(x) + [IN](shortcut)) with act = LeakyReLU(slope) or identity (slope=None).
    models/blocks.py:456-463 (+ :501, :590, :662, :678).  emit_split: also write the bf16 (hi, lo)
    planes of the result (attached as ``out._pcrcg_split``) for a following :func:`linear`."""
    _need_cuda(x, shortcut)
    if isinstance(x, PlaneTensor):
        x = x.dense()
    n, c = x.shape
    mean, rstd, seg = _stats_of(x, segments, eps)
    x = _f32c(x)
    scm = scr = None
    sc_planes = None
    if shortcut is not None:
        if isinstance(shortcut, PlaneTensor):
            if shortcut_norm or c % 4 != 0:
                shortcut = shortcut.dense()
            else:
                sc_planes, shortcut = shortcut._pcrcg_split, None
        if shortcut is not None:
            if shortcut_norm:
                scm, scr, _ = _stats_of(shortcut, segments, eps)
            shortcut = _f32c(shortcut)
    p = lambda t: t.data_ptr() if t is not None else None
    sp = _split_planes(n, c, x.device) if (emit_split and c % 8 == 0 and not _force_simt) else None
    rp = None
    if emit_rowpos and sp is not None and c % 4 == 0 and ((c // 4) & (c // 4 - 1)) == 0:
        rp = torch.empty(n, dtype=torch.uint8, device=x.device)       # (row sum > 0) for the KPConv neighbour count
    # planes_only: the consumer is a tensor-core contraction (c % 64 == 0 covers KPConv's plane kernel and the Linear)
    planes_only = bool(planes_only and sp is not None and c % 64 == 0 and n >= 1 and (rp is not None or not emit_rowpos))
    out = None if planes_only else torch.empty_like(x)
    with torch.cuda.device(x.device):
        if sc_planes is not None:
            check(lib().pcrcg_norm_act_planes_dev(x.data_ptr(), n, c, seg.data_ptr(), seg.shape[0] - 1, mean.data_ptr(), rstd.data_ptr(),
                                                  sc_planes[0].data_ptr(), sc_planes[1].data_ptr(), sc_planes[2], None, None,
                                                  -1.0 if slope is None else float(slope), p(out),
                                                  sp[0].data_ptr() if sp else None, sp[1].data_ptr() if sp else None, sp[2] if sp else 0,
                                                  rp.data_ptr() if rp is not None else None, _stream()))
        else:
            check(lib().pcrcg_norm_act_dev(x.data_ptr(), n, c, seg.data_ptr(), seg.shape[0] - 1, mean.data_ptr(), rstd.data_ptr(),
                                           p(shortcut), p(scm), p(scr), -1.0 if slope is None else float(slope), p(out),
                                           sp[0].data_ptr() if sp else None, sp[1].data_ptr() if sp else None, sp[2] if sp else 0,
                                           rp.data_ptr() if rp is not None else None, _stream()))
    if planes_only:
        return PlaneTensor(sp[0], sp[1], sp[2], n, c, rp)
    if sp is not None:
        _attach(out, "_pcrcg_split", sp)
    if rp is not None:
        _attach(out, "_pcrcg_rowpos", rp)
    return out


def add_act(x, shortcut, slope):
    """LeakyReLU(x + shortcut) without normalisation."""
    x, shortcut = _f32c(x), _f32c(shortcut)
    n, c = x.shape
    seg = _seg_starts(n, None, x.device)
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        check(lib().pcrcg_norm_act_dev(x.data_ptr(), n, c, seg.data_ptr(), 1, None, None, shortcut.data_ptr(), None, None,
                                       float(slope), out.data_ptr(), None, None, 0, None, _stream()))
    return out


def max_pool(x, inds):
    """models/blocks.py:86-102.  Features held as bf16 planes (:class:`PlaneTensor`) are pooled as planes."""
    _need_cuda(x, inds)
    if isinstance(x, PlaneTensor):
        hi, lo, ld = x._pcrcg_split
        idx, is64, H, stride = _idx(inds)
        nq, c = idx.shape[0], x.shape[1]
        if c % 4 == 0:
            oh, ol, ldo = _split_planes(nq, c, hi.device)
            with torch.cuda.device(hi.device):
                check(lib().pcrcg_max_pool_planes_dev(hi.data_ptr(), lo.data_ptr(), x.shape[0], c, ld, idx.data_ptr(), is64, nq, H, stride,
                                                      oh.data_ptr(), ol.data_ptr(), ldo, _stream()))
            return PlaneTensor(oh, ol, ldo, nq, c)
        x = x.dense()
    x = _f32c(x)
    idx, is64, H, stride = _idx(inds)
    nq = idx.shape[0]
    out = torch.empty((nq, x.shape[1]), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        check(lib().pcrcg_max_pool_dev(x.data_ptr(), x.shape[0], x.shape[1], idx.data_ptr(), is64, nq, H, stride, out.data_ptr(), _stream()))
    return out


def closest_pool(x, inds):
    """models/blocks.py:71-83"""
    _need_cuda(x, inds)
    x = _f32c(x)
    idx, is64, H, stride = _idx(inds)
    nq = idx.shape[0]
    out = torch.empty((nq, x.shape[1]), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        check(lib().pcrcg_closest_pool_dev(x.data_ptr(), x.shape[0], x.shape[1], idx.data_ptr(), is64, nq, stride, out.data_ptr(), _stream()))
    return out


def descriptor_head(x, final_feats_dim):
    """models/architectures.py:572-582: -> (feats_f [N,F] L2-normalised, scores_overlap [N], scores_saliency [N])"""
    _need_cuda(x)
    x = _f32c(x)
    n, c = x.shape
    F = int(final_feats_dim)
    if c != F + 2:
        raise RuntimeError("descriptor_head: expected final_feats_dim + 2 columns")
    feats = torch.empty((n, F), dtype=torch.float32, device=x.device)
    ov = torch.empty(n, dtype=torch.float32, device=x.device)
    sa = torch.empty(n, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        check(lib().pcrcg_descriptor_head_dev(x.data_ptr(), n, F, feats.data_ptr(), ov.data_ptr(), sa.data_ptr(), _stream()))
    return feats, ov, sa


def force_simt_contraction(on):
    """True: fp32 CUDA-core contraction (parity anchor); False: tcgen05 tensor cores where shapes allow."""
    global _force_simt
    _force_simt = bool(on)
    lib().pcrcg_gemm_force_simt(1 if on else 0)


# =================================================================================================
# Bottleneck GNN operators (models/gcn.py, models/architectures.py:528-565)
# =================================================================================================
def group_starts(lens, group=1):
    """lens [B] int32 cuda -> int32 [B // group + 1] row starts of groups of `group` consecutive clouds (device, no sync)."""
    lens = _i32c(lens)
    nb = lens.shape[0]
    out = torch.empty(nb // group + 1, dtype=torch.int32, device=lens.device)
    with torch.cuda.device(lens.device):
        check(lib().pcrcg_group_starts_dev(lens.data_ptr(), nb, int(group), out.data_ptr(), None, _stream()))
    return out


def cloud_starts(lens):
    """lens [B] (any int tensor / sequence) -> int32 device-agnostic row starts [B+1] (host computation: B is tiny)."""
    l = torch.as_tensor(lens, dtype=torch.int64).cpu()
    return torch.cat([torch.zeros(1, dtype=torch.int64), torch.cumsum(l, 0)]).to(torch.int32)


def gemm(a, b, b_is_nk, out=None):
    """a [M,K] @ (b [N,K]^T if b_is_nk else b [K,N]) -> [M,N].  2-D fp32 CUDA tensors whose LAST stride is 1; row strides
    are passed through (column slices of wider matrices -- attention heads -- need no copy)."""
    _need_cuda(a, b)
    assert a.dtype == torch.float32 and b.dtype == torch.float32 and a.stride(1) == 1 and b.stride(1) == 1
    m, k = a.shape
    n = b.shape[0] if b_is_nk else b.shape[1]
    if out is None:
        out = torch.empty((m, n), dtype=torch.float32, device=a.device)
    assert out.shape == (m, n) and out.stride(1) == 1
    with torch.cuda.device(a.device):
        check(lib().pcrcg_gemm_dev(a.data_ptr(), a.stride(0), b.data_ptr(), b.stride(0), 1 if b_is_nk else 0, out.data_ptr(), out.stride(0),
                                   m, n, k, None, _stream()))
    return out


def knn(points, starts, k):
    """models/gcn.py:48-51.  points [N,3], starts = cloud row starts [B+1] -> int32 [N,k] global rows."""
    _need_cuda(points)
    points = _f32c(points)
    starts = _i32c(starts.to(points.device))
    out = torch.empty((points.shape[0], k), dtype=torch.int32, device=points.device)
    with torch.cuda.device(points.device):
        check(lib().pcrcg_knn_dev(points.data_ptr(), points.shape[0], starts.data_ptr(), starts.shape[0] - 1, int(k), out.data_ptr(), _stream()))
    return out


def edge_conv_max(uv, cout, knn_idx, starts, slope=0.2, eps=1e-5):
    """uv [N, 2*cout] = (u | v) node-level halves of the edge convolution -> act(IN2d(u_n + v_j)) max-reduced over the k
    edges [N, cout] (models/gcn.py:125-131), statistics per cloud."""
    _need_cuda(uv, knn_idx)
    n, k = knn_idx.shape
    dev = uv.device
    starts = _i32c(starts.to(dev))
    nb = starts.shape[0] - 1
    m = torch.empty((n, cout), dtype=torch.float32, device=dev)
    acc = torch.empty((nb, 2, cout), dtype=torch.float64, device=dev)          # cleared by pcrcg_edge_max_stats_dev
    with torch.cuda.device(dev):
        check(lib().pcrcg_edge_max_stats_dev(uv.data_ptr(), uv.stride(0), uv.data_ptr() + 4 * cout, uv.stride(0), knn_idx.data_ptr(), n, cout, k,
                                             starts.data_ptr(), nb, m.data_ptr(), acc.data_ptr(), _stream()))
        _stats_end(m, starts, acc, eps)
    return instance_norm_act(m, starts, slope, eps=eps)


def bias_act(x, bias, slope=None, out=None):
    """act(x + bias): Conv1d bias, slope None = identity, 0.0 = ReLU.  In place when out is x."""
    _need_cuda(x, bias)
    x = _f32c(x)
    n, c = x.shape
    out = torch.empty_like(x) if out is None else out
    _strip_derived(out)                                   # in place: planes / statistics of the old values are stale
    with torch.cuda.device(x.device):
        check(lib().pcrcg_bias_act_dev(x.data_ptr(), n, c, bias.data_ptr() if bias is not None else None,
                                       -1.0 if slope is None else float(slope), out.data_ptr(), _stream()))
    return out


def softmax_rows_(x, scale=1.0):
    """in place: x <- softmax(scale * x, dim=1)"""
    _need_cuda(x)
    assert x.dtype == torch.float32 and x.dim() == 2 and x.stride(1) == 1
    _strip_derived(x)
    with torch.cuda.device(x.device):
        check(lib().pcrcg_softmax_rows_dev(x.data_ptr(), x.shape[0], x.shape[1], x.stride(0), float(scale), _stream()))
    return x


def l2_normalize(x, eps=1e-12):
    _need_cuda(x)
    x = _f32c(x)
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        check(lib().pcrcg_l2norm_rows_dev(x.data_ptr(), x.shape[0], x.shape[1], float(eps), out.data_ptr(), _stream()))
    return out
