"""Colour path: ``projection.py`` of the reference (class ``Projection``) and the un-projection of
2D image features onto the 3D points (``models/architectures.py:273-307,360-370``), on the GPU.

``Projection(intrinsic_matrix, thresh).projection(points, depth_map, world2camera)`` keeps the
reference's signature and return values (``inds2d`` LongTensor [M,2] in (x, y) order, ``inds3d``
LongTensor [M], ascending), returned on the device the points came from.
"""
import ctypes as C

import numpy as np
import torch

from ._lib import lib, check
from .ops import _f32c, _stream, _ws


MAX_VIEWS = 8                     # csrc/projection.cu: ViewSet


def _mat16(m):
    m = torch.as_tensor(m, dtype=torch.float32).detach().cpu()
    if tuple(m.shape) == (3, 3):                      # projection.py:22-25
        e = torch.eye(4)
        e[:3, :3] = m
        m = e
    return np.ascontiguousarray(m.numpy().reshape(16), dtype=np.float32)


class Projection(object):
    def __init__(self, intrinsic_matrix=0, thresh=0.1, device="cuda"):
        self.intrinsics = intrinsic_matrix
        self.thresh = thresh
        self.device = torch.device(device)

    def projection(self, points, depth_map, world2camera):
        src_dev = points.device if torch.is_tensor(points) else torch.device("cpu")
        dev = src_dev if src_dev.type == "cuda" else self.device
        pts = _f32c(torch.as_tensor(points).to(dev))
        depth = _f32c(torch.as_tensor(depth_map).to(dev))
        depth = depth.reshape(depth.shape[-2], depth.shape[-1])          # projection.py:41 squeeze(0)
        H, W = depth.shape
        n = pts.shape[0]
        k4, w2c = _mat16(self.intrinsics), _mat16(world2camera)
        L = lib()
        with torch.cuda.device(dev):
            i2 = torch.empty((max(n, 1), 2), dtype=torch.int64, device=dev)
            i3 = torch.empty(max(n, 1), dtype=torch.int64, device=dev)
            cnt = torch.zeros(1, dtype=torch.int32, device=dev)
            ws = _ws(L.pcrcg_projection_ws_bytes(n), dev)
            check(L.pcrcg_projection_dev(pts.data_ptr(), n, depth.data_ptr(), H, W, w2c.ctypes.data, k4.ctypes.data, float(self.thresh),
                                         i2.data_ptr(), i3.data_ptr(), cnt.data_ptr(), ws.data_ptr(), ws.numel(), _stream()))
            m = int(cnt.item())
        return i2[:m].to(src_dev), i3[:m].to(src_dev)


def unproject_features(points, views, base=None, thresh=0.1):
    """Fused projection + feature gather + scatter.

    points [N,3] cuda; views: list in the reference's WRITE order (e.g. src image 2, src image 1,
    tgt image 2, tgt image 1) of dicts with ``depth`` [H,W], ``world2camera`` [4,4], ``intrinsics``
    [4,4], ``feature2d`` [C,H,W] (cuda), optional ``valid_map`` [H,W] (the reference batch's [W,H] layout is accepted
    and transposed; a square map is taken as [H,W]), and ``rows`` = (lo, hi): the range of point rows (the cloud) the
    view belongs to.  All views share one (C, H, W); at most 8 views per call.  Returns x [N, C+1]."""
    pts = _f32c(points)
    dev = pts.device
    n = pts.shape[0]
    nv = len(views)
    if not 1 <= nv <= MAX_VIEWS:
        raise RuntimeError(f"unproject_features: between 1 and {MAX_VIEWS} views per call (got {nv})")
    keep = []
    dptr, fptr, vptr = (C.c_void_p * nv)(), (C.c_void_p * nv)(), (C.c_void_p * nv)()
    w2c = np.zeros((nv, 16), np.float32)
    k4 = np.zeros((nv, 16), np.float32)
    lo, hi = np.zeros(nv, np.int32), np.zeros(nv, np.int32)
    Cc = H = W = None
    for v, view in enumerate(views):
        d = _f32c(torch.as_tensor(view["depth"]).to(dev))
        d = d.reshape(d.shape[-2], d.shape[-1])
        f = _f32c(torch.as_tensor(view["feature2d"]).to(dev))
        if f.dim() != 3:
            raise RuntimeError(f"unproject_features: view {v}: feature2d must be [C,H,W], got {tuple(f.shape)}")
        if Cc is None:
            Cc, H, W = f.shape
        # the kernel indexes every view with ONE (C, H, W): a view of another size would read out of bounds
        if tuple(f.shape) != (Cc, H, W):
            raise RuntimeError(f"unproject_features: view {v}: feature2d {tuple(f.shape)} differs from view 0's {(Cc, H, W)}")
        if tuple(d.shape) != (H, W):
            raise RuntimeError(f"unproject_features: view {v}: depth {tuple(d.shape)} does not match feature2d's (H, W) = {(H, W)}")
        vm = view.get("valid_map")
        if vm is not None:
            vm = torch.as_tensor(vm).to(dev)
            vm = vm.reshape(vm.shape[-2], vm.shape[-1])
            if tuple(vm.shape) == (W, H) and H != W:
                # the reference batch stores valid maps as [W,H] and transposes them in KPFCNN.forward
                # (models/architectures.py:287-307): accept that layout as is
                vm = vm.t()
            if tuple(vm.shape) != (H, W):
                raise RuntimeError(f"unproject_features: view {v}: valid_map {tuple(vm.shape)} is neither (H, W) = {(H, W)} nor (W, H)")
            vm = _f32c(vm)
        keep += [d, f, vm]
        dptr[v], fptr[v], vptr[v] = d.data_ptr(), f.data_ptr(), (vm.data_ptr() if vm is not None else None)
        w2c[v], k4[v] = _mat16(view["world2camera"]), _mat16(view["intrinsics"])
        lo[v], hi[v] = view.get("rows", (0, n))
        if not 0 <= lo[v] <= hi[v] <= n:
            raise RuntimeError(f"unproject_features: view {v}: rows {(int(lo[v]), int(hi[v]))} outside [0, {n}]")
    out = torch.empty((n, Cc + 1), dtype=torch.float32, device=dev)
    b = _f32c(base.reshape(-1)) if base is not None else None
    with torch.cuda.device(dev):
        check(lib().pcrcg_project_scatter_dev(pts.data_ptr(), n, nv, dptr, fptr, vptr, w2c.ctypes.data, k4.ctypes.data, lo.ctypes.data,
                                              hi.ctypes.data, H, W, Cc, float(thresh), b.data_ptr() if b is not None else None,
                                              out.data_ptr(), _stream()))
    return out


VIEW_DTYPE = np.dtype([("depth", np.uint64), ("feat", np.uint64), ("valid", np.uint64), ("w2c", np.float32, 12), ("k4", np.float32, 12)])


def unproject_features_batch(points, lengths, views_per_cloud, base=None, thresh=0.1):
    """:func:`unproject_features` for a STACKED batch: points [N,3] cuda = clouds of ``lengths`` [B] rows each;
    ``views_per_cloud[c]`` = the views of cloud c in the reference's write order (image 2, then image 1:
    models/architectures.py:367-370), dicts as in :func:`unproject_features` (``rows`` is implied by the cloud).  One launch for
    any number of clouds and views.  Returns x [N, C+1]."""
    pts = _f32c(points)
    dev = pts.device
    n = pts.shape[0]
    lens = np.asarray(torch.as_tensor(lengths).cpu(), dtype=np.int64).reshape(-1)
    nb = len(lens)
    if len(views_per_cloud) != nb or int(lens.sum()) != n:
        raise RuntimeError("unproject_features_batch: one view list per cloud, lengths summing to the number of points")
    cloud_starts = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    view_starts = np.concatenate([[0], np.cumsum([len(v) for v in views_per_cloud])]).astype(np.int32)
    nv = int(view_starts[-1])
    if nv == 0:
        raise RuntimeError("unproject_features_batch: no views")
    rec = np.zeros(nv, VIEW_DTYPE)
    keep, Cc, H, W, k = [], None, None, None, 0
    for c, vlist in enumerate(views_per_cloud):
        for view in vlist:
            d = _f32c(torch.as_tensor(view["depth"]).to(dev))
            d = d.reshape(d.shape[-2], d.shape[-1])
            f = _f32c(torch.as_tensor(view["feature2d"]).to(dev))
            if f.dim() != 3:
                raise RuntimeError(f"unproject_features_batch: cloud {c}: feature2d must be [C,H,W], got {tuple(f.shape)}")
            if Cc is None:
                Cc, H, W = f.shape
            if tuple(f.shape) != (Cc, H, W) or tuple(d.shape) != (H, W):
                raise RuntimeError(f"unproject_features_batch: cloud {c}: feature2d {tuple(f.shape)} / depth {tuple(d.shape)} differ from "
                                   f"the first view's (C, H, W) = {(Cc, H, W)}")
            vm = view.get("valid_map")
            if vm is not None:
                vm = torch.as_tensor(vm).to(dev)
                vm = vm.reshape(vm.shape[-2], vm.shape[-1])
                if tuple(vm.shape) == (W, H) and H != W:
                    vm = vm.t()
                if tuple(vm.shape) != (H, W):
                    raise RuntimeError(f"unproject_features_batch: cloud {c}: valid_map {tuple(vm.shape)} is neither (H, W) nor (W, H)")
                vm = _f32c(vm)
            keep += [d, f, vm]
            rec[k]["depth"], rec[k]["feat"], rec[k]["valid"] = d.data_ptr(), f.data_ptr(), (vm.data_ptr() if vm is not None else 0)
            rec[k]["w2c"], rec[k]["k4"] = _mat16(view["world2camera"])[:12], _mat16(view["intrinsics"])[:12]
            k += 1
    out = torch.empty((n, Cc + 1), dtype=torch.float32, device=dev)
    b = _f32c(base.reshape(-1)) if base is not None else None
    cs = torch.from_numpy(cloud_starts).to(dev)
    vs = torch.from_numpy(view_starts).to(dev)
    vd = torch.from_numpy(rec.view(np.uint8).reshape(-1).copy()).to(dev)
    with torch.cuda.device(dev):
        check(lib().pcrcg_project_scatter_batch_dev(pts.data_ptr(), n, cs.data_ptr(), nb, vs.data_ptr(), vd.data_ptr(), H, W, Cc, float(thresh),
                                                    b.data_ptr() if b is not None else None, out.data_ptr(), _stream()))
    return out
