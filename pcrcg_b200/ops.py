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
        m = int(out_lens.sum().item())
    return out[:m], out_lens


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
        """-> rows [Nq,width] i32 (None if width == 0), counts [Nq] i32, max_count [1] i32 (device)."""
        _need_cuda(queries, q_lens)
        queries, q_lens = _f32c(queries), _i32c(q_lens)
        nq = queries.shape[0]
        dev = queries.device
        L = lib()
        with torch.cuda.device(dev):
            rows = torch.empty((nq, width), dtype=torch.int32, device=dev) if width > 0 else None
            counts = torch.empty(nq, dtype=torch.int32, device=dev) if want_counts else None
            mx = torch.zeros(1, dtype=torch.int32, device=dev)
            check(L.pcrcg_radius_query_dev(queries.data_ptr(), nq, q_lens.data_ptr(), self.ns, self.nb, self.radius,
                                           int(width), int(width), rows.data_ptr() if rows is not None else None,
                                           counts.data_ptr() if counts is not None else None, mx.data_ptr(),
                                           self.ws.data_ptr(), self.ws.numel(), _stream()))
        return rows, counts, mx


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
