"""Synthetic inputs shaped like the reference's datasets.

Pure NumPy data generation -- no compute of the hot path happens here.  Every generator is
deterministic in its ``seed``.  Shapes follow the reference's loaders:

* 3DMatch / 3DLoMatch fragments: ``datasets/indoor.py:123-147`` (fragments pre-voxelised at 2.5 cm,
  capped at 30 000 points, float32 [N,3]),
* KITTI scans: ``datasets/kitti.py:90-182`` (64-beam scan, voxelised at ``first_subsampling_dl``),
* RGB-D views of the colour path: ``datasets/indoor.py:468-630`` (depth [120,160] in metres,
  4x4 intrinsics from ``datasets/visualize.py:244-276``, valid maps, world->camera poses).
"""
import math

import numpy as np


# ------------------------------------------------------------------------------------------------
# small SE(3) helpers
def _rot_axis_angle(axis, ang):
    axis = np.asarray(axis, np.float64)
    axis = axis / np.linalg.norm(axis)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + math.sin(ang) * K + (1 - math.cos(ang)) * (K @ K)


def random_se3(rng, max_angle_deg=60.0, max_trans=1.0):
    axis = rng.normal(size=3)
    ang = math.radians(rng.uniform(0, max_angle_deg))
    R = _rot_axis_angle(axis, ang)
    t = rng.normal(size=3)
    t = t / np.linalg.norm(t) * rng.uniform(0, max_trans)
    T = np.eye(4)
    T[:3, :3], T[:3, 3] = R, t
    return T


# ------------------------------------------------------------------------------------------------
# indoor scene: room shell + boxes + blobs, sampled uniformly on the surfaces
def _sample_rect(rng, origin, u, v, density):
    area = np.linalg.norm(np.cross(u, v))
    n = max(int(area * density), 1)
    ab = rng.random((n, 2))
    return origin[None] + ab[:, :1] * u[None] + ab[:, 1:] * v[None]


def _sample_box(rng, centre, size, density, R=None):
    pts = []
    sx, sy, sz = size
    corners = [
        ((-sx, -sy, -sz), (2 * sx, 0, 0), (0, 2 * sy, 0)), ((-sx, -sy, sz), (2 * sx, 0, 0), (0, 2 * sy, 0)),
        ((-sx, -sy, -sz), (2 * sx, 0, 0), (0, 0, 2 * sz)), ((-sx, sy, -sz), (2 * sx, 0, 0), (0, 0, 2 * sz)),
        ((-sx, -sy, -sz), (0, 2 * sy, 0), (0, 0, 2 * sz)), ((sx, -sy, -sz), (0, 2 * sy, 0), (0, 0, 2 * sz)),
    ]
    for o, u, v in corners:
        pts.append(_sample_rect(rng, np.array(o, float) / 2, np.array(u, float) / 2, np.array(v, float) / 2, density))
    p = np.concatenate(pts)
    if R is not None:
        p = p @ R.T
    return p + np.asarray(centre)[None]


def _sample_ellipsoid(rng, centre, radii, density):
    a, b, c = radii
    area = 4 * math.pi * (((a * b) ** 1.6 + (a * c) ** 1.6 + (b * c) ** 1.6) / 3) ** (1 / 1.6)
    n = max(int(area * density), 1)
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return d * np.array(radii)[None] + np.asarray(centre)[None]


def indoor_scene(rng, density=2500.0, room=(3.0, 2.5, 2.5)):
    """Surface samples (float64 [n,3]) of a furnished room; ``density`` in points per square metre."""
    X, Y, Z = (room[0] * rng.uniform(0.9, 1.6), room[1] * rng.uniform(0.9, 1.6), room[2])
    parts = [
        _sample_rect(rng, np.array([0, 0, 0.0]), np.array([X, 0, 0.0]), np.array([0, Y, 0.0]), density),      # floor
        _sample_rect(rng, np.array([0, 0, 0.0]), np.array([X, 0, 0.0]), np.array([0, 0, Z]), density),         # wall y=0
        _sample_rect(rng, np.array([0, 0, 0.0]), np.array([0, Y, 0.0]), np.array([0, 0, Z]), density),         # wall x=0
        _sample_rect(rng, np.array([X, 0, 0.0]), np.array([0, Y, 0.0]), np.array([0, 0, Z]), density),         # wall x=X
        _sample_rect(rng, np.array([0, Y, 0.0]), np.array([X, 0, 0.0]), np.array([0, 0, Z]), density),         # wall y=Y
    ]
    for _ in range(int(rng.integers(3, 7))):
        size = rng.uniform(0.2, 0.9, size=3)
        c = np.array([rng.uniform(0.3, X - 0.3), rng.uniform(0.3, Y - 0.3), size[2] / 2])
        parts.append(_sample_box(rng, c, size, density, _rot_axis_angle([0, 0, 1], rng.uniform(0, math.pi))))
    for _ in range(int(rng.integers(2, 5))):
        r = rng.uniform(0.1, 0.35, size=3)
        c = np.array([rng.uniform(0.3, X - 0.3), rng.uniform(0.3, Y - 0.3), rng.uniform(0.2, 1.4)])
        parts.append(_sample_ellipsoid(rng, c, r, density))
    return np.concatenate(parts), (X, Y, Z)


def _frustum_crop(pts, cam_pos, look_at, fov_deg=(75.0, 60.0), zmax=3.5):
    f = look_at - cam_pos
    f /= np.linalg.norm(f)
    up = np.array([0, 0, 1.0])
    r = np.cross(f, up)
    r /= np.linalg.norm(r)
    u = np.cross(r, f)
    d = pts - cam_pos[None]
    z = d @ f
    x = d @ r
    y = d @ u
    tx, ty = math.tan(math.radians(fov_deg[0] / 2)), math.tan(math.radians(fov_deg[1] / 2))
    return (z > 0.3) & (z < zmax) & (np.abs(x) < tx * z) & (np.abs(y) < ty * z)


def voxel_downsample_np(pts, dl):
    """NumPy voxel-grid mean (data generation only -- NOT the reference's barycentre order)."""
    k = np.floor(pts / dl).astype(np.int64)
    k -= k.min(0)
    dims = k.max(0) + 1
    key = (k[:, 0] * dims[1] + k[:, 1]) * dims[2] + k[:, 2]
    order = np.argsort(key, kind="stable")
    key_s = key[order]
    starts = np.flatnonzero(np.r_[True, key_s[1:] != key_s[:-1]])
    sums = np.add.reduceat(pts[order], starts, axis=0)
    cnt = np.diff(np.r_[starts, len(key_s)])
    return sums / cnt[:, None]


def match3d_pair(seed, n_target=20000, overlap="high", dl=0.025, lattice=0.0, noise=0.002, max_points=30000):
    """One 3DMatch-shaped fragment pair -> (src [Ns,3] f32, tgt [Nt,3] f32, T_src_to_tgt [4,4] f64).

    overlap='high' (3DMatch-like, >=30 % shared view) or 'low' (3DLoMatch-like, 10-30 %).
    lattice>0 rounds coordinates to that pitch (real 3DMatch fragments sit on a ~6 mm TSDF lattice,
    which is what produces equal-distance ties in the neighbour lists)."""
    rng = np.random.default_rng(seed)
    scene, (X, Y, Z) = indoor_scene(rng, density=3.2 / (dl * dl))
    scene = scene + rng.normal(scale=noise, size=scene.shape)
    centre = np.array([X / 2, Y / 2, 1.0])
    frags = []
    base_ang = rng.uniform(0, 2 * math.pi)
    dang = rng.uniform(0.25, 0.7) if overlap == "high" else rng.uniform(1.7, 2.3)
    for k in range(2):
        ang = base_ang + k * dang
        cam = centre + np.array([0.35 * X * math.cos(ang + math.pi), 0.35 * Y * math.sin(ang + math.pi), 0.5])
        look = centre + np.array([0.5 * X * math.cos(ang), 0.5 * Y * math.sin(ang), -0.3])
        m = _frustum_crop(scene, cam, look)
        p = voxel_downsample_np(scene[m], dl)
        # scale the crop towards the target size by trimming depth
        n_keep = int(n_target * rng.uniform(0.75, 1.2))
        if len(p) > n_keep:
            d = np.linalg.norm(p - cam[None], axis=1)
            p = p[d <= np.partition(d, n_keep)[n_keep]]
        if len(p) > max_points:
            p = p[rng.permutation(len(p))[:max_points]]             # datasets/indoor.py:142-147
        frags.append(p[rng.permutation(len(p))])
    T = random_se3(rng, 60.0, 1.0)
    src = (frags[0] - centre) @ np.linalg.inv(T[:3, :3]).T          # src lives in its own frame
    src = src - np.linalg.inv(T[:3, :3]) @ T[:3, 3]
    tgt = frags[1] - centre
    if lattice > 0:
        src = np.round(src / lattice) * lattice
        tgt = np.round(tgt / lattice) * lattice
    return src.astype(np.float32), tgt.astype(np.float32), T


# ------------------------------------------------------------------------------------------------
# KITTI-shaped scan: 64 beams x ~1900 azimuths ray-cast onto ground + street walls + boxes
def kitti_scan(seed, n_beams=64, n_az=1900, noise=0.02, max_range=80.0):
    rng = np.random.default_rng(seed)
    elev = np.radians(np.linspace(-24.8, 2.0, n_beams))
    az = np.linspace(-math.pi, math.pi, n_az, endpoint=False) + rng.uniform(0, 1e-3)
    e, a = np.meshgrid(elev, az, indexing="ij")
    d = np.stack([np.cos(e) * np.cos(a), np.cos(e) * np.sin(a), np.sin(e)], -1).reshape(-1, 3)
    o = np.array([0.0, 0.0, 1.73])
    t_best = np.full(len(d), np.inf)
    # ground z = 0
    with np.errstate(divide="ignore", invalid="ignore"):
        t = -o[2] / d[:, 2]
    t_best = np.where((t > 0) & (t < t_best), t, t_best)
    # street walls y = +-w
    for w in (rng.uniform(6, 12), -rng.uniform(6, 12)):
        with np.errstate(divide="ignore", invalid="ignore"):
            t = (w - o[1]) / d[:, 1]
        z = o[2] + t * d[:, 2]
        ok = (t > 0) & (z < rng.uniform(4, 10)) & (z > 0)
        t_best = np.where(ok & (t < t_best), t, t_best)
    # boxes (cars / poles): axis-aligned slabs
    for _ in range(int(rng.integers(12, 30))):
        c = np.array([rng.uniform(-50, 50), rng.uniform(-5.5, 5.5), 0.0])
        s = np.array([rng.uniform(0.3, 4.5), rng.uniform(0.3, 2.0), rng.uniform(1.0, 2.5)])
        lo, hi = c - np.array([s[0] / 2, s[1] / 2, 0]), c + np.array([s[0] / 2, s[1] / 2, s[2]])
        with np.errstate(divide="ignore", invalid="ignore"):
            t1, t2 = (lo[None] - o[None]) / d, (hi[None] - o[None]) / d
        tn = np.nanmax(np.minimum(t1, t2), axis=1)
        tf = np.nanmin(np.maximum(t1, t2), axis=1)
        ok = (tn < tf) & (tn > 0)
        t_best = np.where(ok & (tn < t_best), tn, t_best)
    keep = np.isfinite(t_best) & (t_best < max_range)
    p = o[None] + t_best[keep, None] * d[keep]
    p = p + rng.normal(scale=noise, size=p.shape)
    return p.astype(np.float32)


def kitti_pair(seed, dl=0.3):
    """Two scans ~10 m apart along the street; returns raw scans (float32, ~120k points each).
    The caller voxelises them at ``dl`` (the reference uses open3d there, datasets/kitti.py:137)."""
    a = kitti_scan(seed * 2 + 0)
    b = kitti_scan(seed * 2 + 0)            # same street ...
    rng = np.random.default_rng(seed * 2 + 1)
    T = np.eye(4)
    T[:3, :3] = _rot_axis_angle([0, 0, 1], math.radians(rng.uniform(-8, 8)))
    T[:3, 3] = [rng.uniform(8, 12), rng.uniform(-0.5, 0.5), 0.0]
    b = ((b.astype(np.float64) - T[:3, 3]) @ T[:3, :3]).astype(np.float32)   # ... seen from a moved car
    return a, b, T


# ------------------------------------------------------------------------------------------------
# RGB-D views for the colour path
def adjust_intrinsic(K, dim_before=(640, 480), dim_after=(160, 120)):
    """Arithmetic of datasets/visualize.py:244-276 (restated; returns a 4x4 float32)."""
    K4 = np.eye(4, dtype=np.float64)
    K4[:3, :3] = np.asarray(K, np.float64)[:3, :3]
    hr, wr = dim_after[1] / dim_before[1], dim_after[0] / dim_before[0]
    if wr >= hr:
        rh, rw = dim_after[1], hr * dim_before[0]
    else:
        rw, rh = dim_after[0], wr * dim_before[1]
    K4[0, 0] *= rw / dim_before[0]
    K4[1, 1] *= rh / dim_before[1]
    K4[0, 2] *= (rw - 1) / (dim_before[0] - 1)
    K4[1, 2] *= (rh - 1) / (dim_before[1] - 1)
    return K4.astype(np.float32)


def render_depth(points, world2camera, K4, hw=(120, 160)):
    """z-buffer splat of ``points`` -> depth [H,W] float32 metres (0 = no hit, like a depth png)."""
    H, W = hw
    cam = points.astype(np.float64) @ world2camera[:3, :3].T + world2camera[:3, 3]
    z = cam[:, 2]
    ok = z > 0.2
    u = (K4[0, 0] * cam[:, 0] / np.where(ok, z, 1) + K4[0, 2])
    v = (K4[1, 1] * cam[:, 1] / np.where(ok, z, 1) + K4[1, 2])
    ui, vi = np.floor(u).astype(np.int64), np.floor(v).astype(np.int64)
    ok &= (ui >= 0) & (ui < W) & (vi >= 0) & (vi < H)
    depth = np.full(H * W, np.inf)
    np.minimum.at(depth, vi[ok] * W + ui[ok], z[ok])
    depth[~np.isfinite(depth)] = 0.0
    return depth.reshape(H, W).astype(np.float32)


def rgbd_views(points, seed, n_views=2, channels=128, hw=(120, 160)):
    """Per cloud: n_views of (depth [H,W], world2camera [4,4], K4 [4,4], feature2d [C,H,W],
    valid_map [H,W]).  feature2d ~ N(0,1) stands in for the Res50UNet output
    (models/resunet.py:163-188), valid_map in [0,1] for the SuperGlue confidence map."""
    rng = np.random.default_rng(seed)
    K4 = adjust_intrinsic(np.array([[585.0, 0, 320.0], [0, 585.0, 240.0], [0, 0, 1.0]]))
    c = points.mean(0).astype(np.float64)
    ext = float(np.linalg.norm(points.std(0))) + 0.5
    views = []
    for _ in range(n_views):
        dirn = rng.normal(size=3)
        dirn[2] = abs(dirn[2]) * 0.3
        dirn /= np.linalg.norm(dirn)
        cam_pos = c + dirn * ext * 1.2
        f = c - cam_pos
        f /= np.linalg.norm(f)
        r = np.cross(f, [0, 0, 1.0])
        r /= np.linalg.norm(r)
        d = np.cross(f, r)
        Rcw = np.stack([r, d, f], 0)                    # camera axes: x right, y down, z forward
        W2C = np.eye(4)
        W2C[:3, :3], W2C[:3, 3] = Rcw, -Rcw @ cam_pos
        depth = render_depth(points, W2C, K4, hw)
        feat = rng.normal(size=(channels,) + hw).astype(np.float32)
        valid = rng.random(hw).astype(np.float32)
        views.append(dict(depth=depth, world2camera=W2C.astype(np.float32), intrinsics=K4, feature2d=feat, valid_map=valid))
    return views
