"""oracle/blocks_port.py -- TEST INFRASTRUCTURE ONLY.

Plain PyTorch fp32 (CPU) restatement of the reference's KPConv operator library, written as pure
functions over explicit tensors.  It is the checker the CUDA kernels are compared with and the
"port" CPU baseline of bench.py; nothing under pcrcg_b200/ imports it.

Pinned by tests/golden/blocks_*.npz, which were produced by importing the reference's own
``models/blocks.py`` in the build container (tests/golden/make_golden.py).

Reference lines followed (relative to /root/reference):
  models/blocks.py:229-374  KPConv.forward (rigid, KP_influence='linear', aggregation 'sum')
  models/blocks.py:433-465  BatchNormBlock (= InstanceNorm1d over all rows, no affine) / bias
  models/blocks.py:473-501  UnaryBlock
  models/blocks.py:578-590  SimpleBlock.forward
  models/blocks.py:650-678  ResnetBottleneckBlock.forward
  models/blocks.py:71-102   closest_pool / max_pool
  models/architectures.py:520-524  encoder loop
"""
import torch
import torch.nn.functional as F


def kpconv(q_pts, s_pts, neighb_inds, x, kernel_points, weights, KP_extent):
    """models/blocks.py:265-374, non-deformable branch."""
    neighb_inds = neighb_inds.long()
    s_pad = torch.cat((s_pts, torch.full((1, 3), 1e6, dtype=s_pts.dtype)), 0)        # :269
    rel = s_pad[neighb_inds] - q_pts[:, None, :]                                      # :272-275  [N,H,3]
    diff = rel[:, :, None, :] - kernel_points[None, None, :, :]                       # :285-286  [N,H,K,3]
    sq = (diff ** 2).sum(dim=3)                                                       # :289      [N,H,K]
    w = torch.clamp(1 - torch.sqrt(sq) / KP_extent, min=0.0).transpose(1, 2)          # :328-329  [N,K,H]
    x_pad = torch.cat((x, torch.zeros_like(x[:1])), 0)                                # :348
    nx = x_pad[neighb_inds]                                                           # :351      [N,H,Cin]
    wf = torch.matmul(w, nx)                                                          # :354      [N,K,Cin]
    out = torch.matmul(wf.permute(1, 0, 2), weights).sum(dim=0)                       # :361-366
    cnt = (nx.sum(dim=-1) > 0.0).sum(dim=-1)                                          # :369-370
    cnt = torch.max(cnt, torch.ones_like(cnt))                                        # :371
    return out / cnt[:, None]                                                         # :372


def instance_norm(x, eps=1e-5, segments=None):
    """models/blocks.py:448,456-463: [N,C] -> InstanceNorm1d over N (biased var, no affine).
    ``segments`` (list of row counts) = one normalisation group per fragment pair when several
    pairs are stacked; None = the reference's single group."""
    if segments is None:
        segments = [x.shape[0]]
    outs, i0 = [], 0
    for n in segments:
        xs = x[i0:i0 + n]
        outs.append(F.instance_norm(xs.t()[None], eps=eps)[0].t())
        i0 += n
    return torch.cat(outs, 0)


def leaky(x):
    return F.leaky_relu(x, 0.1)


def unary(x, W, relu=True, segments=None, use_bn=True, bias=None):
    """UnaryBlock, models/blocks.py:496-501.  W is nn.Linear.weight [Cout, Cin]."""
    x = x @ W.t()
    x = instance_norm(x, segments=segments) if use_bn else x + bias
    return leaky(x) if relu else x


def max_pool(x, inds):
    """models/blocks.py:86-102"""
    x_pad = torch.cat((x, torch.zeros_like(x[:1])), 0)
    return x_pad[inds.long()].max(dim=1)[0]


def closest_pool(x, inds):
    """models/blocks.py:71-83"""
    x_pad = torch.cat((x, torch.zeros_like(x[:1])), 0)
    return x_pad[inds[:, 0].long()]


def simple_block(x, q_pts, s_pts, inds, p, segments=None):
    """SimpleBlock.forward, models/blocks.py:578-590.  p: dict(kernel_points, weights, KP_extent)."""
    y = kpconv(q_pts, s_pts, inds, x, p["kernel_points"], p["weights"], p["KP_extent"])
    return leaky(instance_norm(y, segments=segments))


def resnetb_block(x, q_pts, s_pts, inds, p, strided, seg_in=None, seg_out=None):
    """ResnetBottleneckBlock.forward, models/blocks.py:650-678.
    p: dict(unary1 [d/4,din] | None, kernel_points, weights, KP_extent, unary2 [d,d/4],
            shortcut [d,din] | None).  seg_in / seg_out: per-pair row counts at the support /
    query resolution."""
    y = unary(x, p["unary1"], relu=True, segments=seg_in) if p.get("unary1") is not None else x
    y = kpconv(q_pts, s_pts, inds, y, p["kernel_points"], p["weights"], p["KP_extent"])
    y = leaky(instance_norm(y, segments=seg_out))
    y = unary(y, p["unary2"], relu=False, segments=seg_out)
    sc = max_pool(x, inds) if strided else x
    if p.get("shortcut") is not None:
        sc = unary(sc, p["shortcut"], relu=False, segments=seg_out)
    return leaky(y + sc)


def encoder(x, batch, blocks, pair_segments=None):
    """models/architectures.py:520-524 with the per-block wiring of models/blocks.py.
    blocks: list of dict(kind='simple'|'resnetb', strided, layer, params).
    pair_segments: list per layer of per-pair row counts (None = single pair)."""
    outs = []
    for b in blocks:
        l = b["layer"]
        if b["strided"]:
            q, s, idx = batch["points"][l + 1], batch["points"][l], batch["pools"][l]
        else:
            q, s, idx = batch["points"][l], batch["points"][l], batch["neighbors"][l]
        seg_in = None if pair_segments is None else pair_segments[l]
        seg_out = None if pair_segments is None else pair_segments[l + 1 if b["strided"] else l]
        if b["kind"] == "simple":
            x = simple_block(x, q, s, idx, b["params"], segments=seg_out)
        else:
            x = resnetb_block(x, q, s, idx, b["params"], b["strided"], seg_in, seg_out)
        outs.append(x)
    return x, outs


# ---------------------------------------------------------------------------------------------
INDOOR_ARCHITECTURE = ["simple", "resnetb", "resnetb_strided", "resnetb", "resnetb", "resnetb_strided",
                       "resnetb", "resnetb", "resnetb_strided", "resnetb", "resnetb"]   # configs/models.py:2-13 (encoder part)


def encoder_blocks_from_state_dict(sd, architecture=INDOOR_ARCHITECTURE, first_subsampling_dl=0.025, conv_radius=2.5,
                                   KP_extent=2.0, prefix=""):
    """Maps a reference ``KPFCNN.encoder_blocks`` state_dict (models/architectures.py:62-100 naming:
    ``<i>.KPConv.weights``, ``<i>.KPConv.kernel_points``, ``<i>.unary1.mlp.weight`` ...) onto the
    block descriptors used by :func:`encoder`."""
    t = lambda k: torch.as_tensor(sd[prefix + k]).float() if (prefix + k) in sd else None
    blocks, layer, r = [], 0, first_subsampling_dl * conv_radius
    for i, name in enumerate(architecture):
        strided = "strided" in name
        p = dict(kernel_points=t(f"{i}.KPConv.kernel_points"), weights=t(f"{i}.KPConv.weights"),
                 KP_extent=r * KP_extent / conv_radius)
        if name.startswith("simple"):
            blocks.append(dict(kind="simple", strided=strided, layer=layer, params=p))
        else:
            p.update(unary1=t(f"{i}.unary1.mlp.weight"), unary2=t(f"{i}.unary2.mlp.weight"),
                     shortcut=t(f"{i}.unary_shortcut.mlp.weight"))
            blocks.append(dict(kind="resnetb", strided=strided, layer=layer, params=p))
        if strided:
            layer += 1
            r *= 2
    return blocks


def decoder(x, skips, batch, unary_weights, final_feats_dim):
    """models/architectures.py:567-582 for the indoor / kitti architecture tail
    (nearest_upsample, unary, nearest_upsample, unary, nearest_upsample, last_unary).
    unary_weights: [W_unary_1, W_unary_2, W_last]  (nn.Linear weights [out, in])."""
    skips = list(skips)
    n_up = len(unary_weights)
    for i, W in enumerate(unary_weights):
        layer = n_up - i                                   # upsample from `layer` to `layer - 1`
        x = closest_pool(x, batch["upsamples"][layer - 1])  # NearestUpsampleBlock, models/blocks.py:704-705
        x = torch.cat([x, skips.pop()], dim=1)
        x = unary(x, W) if i < n_up - 1 else x @ W.t()      # UnaryBlock / LastUnaryBlock
    feats = F.normalize(x[:, :final_feats_dim], p=2, dim=1)
    so = torch.clamp(torch.sigmoid(x[:, final_feats_dim]), 0, 1)
    ss = torch.clamp(torch.sigmoid(x[:, final_feats_dim + 1]), 0, 1)
    fix = lambda t: torch.where(torch.isfinite(t), t, torch.zeros_like(t))
    return feats, fix(so), fix(ss), x
