"""KPConv operator library with the reference's module API (``models/blocks.py``), computed by the
CUDA kernels of libpcrcg_b200.so.  Forward only (``torch.no_grad`` inference path).

Constructor arguments, attribute and PARAMETER NAMES match the reference so a reference
``state_dict`` loads unchanged (``KPConv.weights``, ``KPConv.kernel_points``, ``unary1.mlp.weight``,
``unary2.mlp.weight``, ``unary_shortcut.mlp.weight``; ``models/blocks.py:175,226,490``).

Differences, all deliberate:
  * ``kernel_points`` are never regenerated with NumPy noise at construction
    (``kernels/kernel_points.py:388-470``): they are a buffer-like Parameter filled from the
    state_dict (or by :func:`set_kernel_points`).
  * only the branches any shipped config uses are implemented (``KP_influence='linear'``,
    ``aggregation_mode='sum'``, rigid kernels); the others raise NotImplementedError.
  * the optional ``batch['pair_segments']`` (list per layer of int32 row starts) normalises each
    fragment pair of a stacked multi-pair batch separately; absent = the reference's behaviour.
"""
import math

import torch
import torch.nn as nn
from torch.nn.parameter import Parameter

from . import ops


def gather(x, idx, method=2):
    """models/blocks.py:27-58 (x[idx]); kept for API completeness."""
    return x[idx]


def closest_pool(x, inds):
    return ops.closest_pool(x, inds)


def max_pool(x, inds):
    return ops.max_pool(x, inds)


def _stat_arg(segments):
    """segments (None = one group) -> the stat_segments argument of the contractions"""
    return True if segments is None else segments


def _segments(batch, layer):
    seg = batch.get("pair_segments") if isinstance(batch, dict) else None
    return None if seg is None else seg[layer]


class KPConv(nn.Module):
    def __init__(self, kernel_size, p_dim, in_channels, out_channels, KP_extent, radius,
                 fixed_kernel_points="center", KP_influence="linear", aggregation_mode="sum",
                 deformable=False, modulated=False):
        super().__init__()
        if deformable or modulated:
            raise NotImplementedError("pcrcg_b200.KPConv: deformable kernels are not used by any PCR-CG config")
        if KP_influence != "linear" or aggregation_mode != "sum":
            raise NotImplementedError("pcrcg_b200.KPConv: only KP_influence='linear', aggregation_mode='sum'")
        if p_dim != 3 or not (1 <= kernel_size <= 15):
            raise NotImplementedError("pcrcg_b200.KPConv: 3D points, at most 15 kernel points")
        self.K, self.p_dim = kernel_size, p_dim
        self.in_channels, self.out_channels = in_channels, out_channels
        self.radius, self.KP_extent = radius, KP_extent
        self.fixed_kernel_points = fixed_kernel_points
        self.KP_influence, self.aggregation_mode = KP_influence, aggregation_mode
        self.deformable, self.modulated = deformable, modulated
        self.weights = Parameter(torch.zeros((self.K, in_channels, out_channels), dtype=torch.float32), requires_grad=False)
        nn.init.kaiming_uniform_(self.weights, a=math.sqrt(5))            # models/blocks.py:207-208
        self.kernel_points = Parameter(torch.zeros((self.K, p_dim), dtype=torch.float32), requires_grad=False)

    def set_kernel_points(self, pts):
        with torch.no_grad():
            self.kernel_points.copy_(torch.as_tensor(pts, dtype=torch.float32))

    def forward(self, q_pts, s_pts, neighb_inds, x, stat_segments=False):
        """stat_segments (extension, default off = the reference signature): the row starts of the normalisation groups of
        the BatchNormBlock that consumes the result; its statistics are then accumulated by the contraction epilogue."""
        return ops.kpconv_forward(q_pts, s_pts, neighb_inds, x, self.kernel_points, self.weights, self.KP_extent, stat_segments)

    def __repr__(self):
        return "KPConv(radius: {:.2f}, extent: {:.2f}, in_feat: {:d}, out_feat: {:d})".format(
            self.radius, self.KP_extent, self.in_channels, self.out_channels)


class BatchNormBlock(nn.Module):
    """models/blocks.py:433-470: InstanceNorm1d over all rows (use_bn) or a bias."""

    def __init__(self, in_dim, use_bn, bn_momentum):
        super().__init__()
        self.bn_momentum, self.use_bn, self.in_dim = bn_momentum, use_bn, in_dim
        if not use_bn:
            self.bias = Parameter(torch.zeros(in_dim, dtype=torch.float32), requires_grad=False)

    def forward(self, x, segments=None, slope=None, emit_split=False, emit_rowpos=False, planes_only=False):
        if self.use_bn:
            return ops.instance_norm_act(x, segments, slope, emit_split=emit_split, emit_rowpos=emit_rowpos, planes_only=planes_only)
        x = x + self.bias
        return x if slope is None else torch.nn.functional.leaky_relu(x, slope)


class _Linear(nn.Module):
    """nn.Linear(bias=False) stand-in with the same parameter name (``weight`` [out,in])."""

    def __init__(self, in_dim, out_dim):
        super().__init__()
        self.weight = Parameter(torch.empty(out_dim, in_dim, dtype=torch.float32), requires_grad=False)
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))

    def forward(self, x, stat_segments=False):
        return ops.linear(x, self.weight, stat_segments)


class UnaryBlock(nn.Module):
    def __init__(self, in_dim, out_dim, use_bn, bn_momentum, no_relu=False):
        super().__init__()
        self.bn_momentum, self.use_bn, self.no_relu = bn_momentum, use_bn, no_relu
        self.in_dim, self.out_dim = in_dim, out_dim
        self.mlp = _Linear(in_dim, out_dim)
        self.batch_norm = BatchNormBlock(out_dim, use_bn, bn_momentum)

    def forward(self, x, batch=None, segments=None, emit_split=False, emit_rowpos=False, planes_only=False):
        y = self.mlp(x, _stat_arg(segments) if self.use_bn else False)
        return self.batch_norm(y, segments, None if self.no_relu else 0.1, emit_split=emit_split, emit_rowpos=emit_rowpos,
                               planes_only=planes_only)


class LastUnaryBlock(nn.Module):
    def __init__(self, in_dim, out_dim, use_bn, bn_momentum, no_relu=False):
        super().__init__()
        self.in_dim, self.out_dim = in_dim, out_dim
        self.mlp = _Linear(in_dim, out_dim)

    def forward(self, x, batch=None):
        return self.mlp(x)


def _block_geometry(block_name, layer_ind, batch):
    """models/blocks.py:580-587 / :652-659"""
    if "strided" in block_name:
        return batch["points"][layer_ind + 1], batch["points"][layer_ind], batch["pools"][layer_ind], layer_ind + 1
    return batch["points"][layer_ind], batch["points"][layer_ind], batch["neighbors"][layer_ind], layer_ind


class SimpleBlock(nn.Module):
    def __init__(self, block_name, in_dim, out_dim, radius, layer_ind, config):
        super().__init__()
        current_extent = radius * config.KP_extent / config.conv_radius
        self.bn_momentum, self.use_bn = config.batch_norm_momentum, config.use_batch_norm
        self.layer_ind, self.block_name, self.in_dim, self.out_dim = layer_ind, block_name, in_dim, out_dim
        self.KPConv = KPConv(config.num_kernel_points, config.in_points_dim, in_dim, out_dim // 2, current_extent, radius,
                             fixed_kernel_points=config.fixed_kernel_points, KP_influence=config.KP_influence,
                             aggregation_mode=config.aggregation_mode, deformable="deform" in block_name,
                             modulated=config.modulated)
        self.batch_norm = BatchNormBlock(out_dim // 2, self.use_bn, self.bn_momentum)

    def forward(self, x, batch):
        q_pts, s_pts, inds, out_layer = _block_geometry(self.block_name, self.layer_ind, batch)
        seg = _segments(batch, out_layer)
        x = self.KPConv(q_pts, s_pts, inds, x, _stat_arg(seg) if self.use_bn else False)
        return self.batch_norm(x, seg, 0.1, emit_split=True)      # feeds the next block's unary1


class ResnetBottleneckBlock(nn.Module):
    def __init__(self, block_name, in_dim, out_dim, radius, layer_ind, config):
        super().__init__()
        current_extent = radius * config.KP_extent / config.conv_radius
        self.bn_momentum, self.use_bn = config.batch_norm_momentum, config.use_batch_norm
        self.block_name, self.layer_ind, self.in_dim, self.out_dim = block_name, layer_ind, in_dim, out_dim
        self.unary1 = UnaryBlock(in_dim, out_dim // 4, self.use_bn, self.bn_momentum) if in_dim != out_dim // 4 else nn.Identity()
        self.KPConv = KPConv(config.num_kernel_points, config.in_points_dim, out_dim // 4, out_dim // 4, current_extent, radius,
                             fixed_kernel_points=config.fixed_kernel_points, KP_influence=config.KP_influence,
                             aggregation_mode=config.aggregation_mode, deformable="deform" in block_name,
                             modulated=config.modulated)
        self.batch_norm_conv = BatchNormBlock(out_dim // 4, self.use_bn, self.bn_momentum)
        self.unary2 = UnaryBlock(out_dim // 4, out_dim, self.use_bn, self.bn_momentum, no_relu=True)
        self.unary_shortcut = (UnaryBlock(in_dim, out_dim, self.use_bn, self.bn_momentum, no_relu=True)
                               if in_dim != out_dim else nn.Identity())

    def forward(self, features, batch):
        q_pts, s_pts, inds, out_layer = _block_geometry(self.block_name, self.layer_ind, batch)
        seg_in, seg_out = _segments(batch, self.layer_ind), _segments(batch, out_layer)
        # unary1's output is gathered by the KPConv aggregation: emit its bf16 planes for the bf16x3 kernel
        # (both intermediates below never leave the block and feed tensor-core contractions only: no fp32 copy is written)
        x = (self.unary1(features, segments=seg_in, emit_split=True, emit_rowpos=True, planes_only=True)
             if isinstance(self.unary1, UnaryBlock) else features)
        stat = _stat_arg(seg_out) if self.use_bn else False
        x = self.KPConv(q_pts, s_pts, inds, x, stat)
        x = self.batch_norm_conv(x, seg_out, 0.1, emit_split=True, planes_only=True)   # feeds unary2
        y = self.unary2.mlp(x, stat)                                 # raw Linear; its norm is fused below
        shortcut = ops.max_pool(features, inds) if "strided" in self.block_name else features
        if isinstance(self.unary_shortcut, UnaryBlock):
            sc_raw = self.unary_shortcut.mlp(shortcut, stat)
            if self.use_bn:
                return ops.instance_norm_act(y, seg_out, 0.1, shortcut=sc_raw, shortcut_norm=True, emit_split=True)
            return ops.add_act(y + self.unary2.batch_norm.bias, sc_raw + self.unary_shortcut.batch_norm.bias, 0.1)
        if self.use_bn:
            return ops.instance_norm_act(y, seg_out, 0.1, shortcut=shortcut, shortcut_norm=False, emit_split=True)
        return ops.add_act(y + self.unary2.batch_norm.bias, shortcut, 0.1)


class NearestUpsampleBlock(nn.Module):
    def __init__(self, layer_ind):
        super().__init__()
        self.layer_ind = layer_ind

    def forward(self, x, batch):
        return closest_pool(x, batch["upsamples"][self.layer_ind - 1])


class MaxPoolBlock(nn.Module):
    def __init__(self, layer_ind):
        super().__init__()
        self.layer_ind = layer_ind

    def forward(self, x, batch):
        return max_pool(x, batch["pools"][self.layer_ind + 1])


def block_decider(block_name, radius, in_dim, out_dim, layer_ind, config):
    """models/blocks.py:387-430"""
    if block_name == "unary":
        return UnaryBlock(in_dim, out_dim, config.use_batch_norm, config.batch_norm_momentum)
    if block_name == "last_unary":
        return LastUnaryBlock(in_dim, config.final_feats_dim + 2, config.use_batch_norm, config.batch_norm_momentum)
    if block_name in ("simple", "simple_strided"):
        return SimpleBlock(block_name, in_dim, out_dim, radius, layer_ind, config)
    if block_name in ("resnetb", "resnetb_strided"):
        return ResnetBottleneckBlock(block_name, in_dim, out_dim, radius, layer_ind, config)
    if block_name in ("max_pool", "max_pool_wide"):
        return MaxPoolBlock(layer_ind)
    if block_name == "nearest_upsample":
        return NearestUpsampleBlock(layer_ind)
    if any(t in block_name for t in ("deformable", "invariant", "equivariant")) or block_name == "global_average":
        raise NotImplementedError("pcrcg_b200: block '%s' is not used by any PCR-CG config" % block_name)
    raise ValueError("Unknown block name in the architecture definition : " + block_name)


class Config(dict):
    """Attribute-access dict with the model keys of configs/*/*.yaml (stand-in for EasyDict)."""
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


ARCHITECTURES = {
    # configs/models.py:2-40 (indoor == kitti)
    "indoor": ["simple", "resnetb", "resnetb_strided", "resnetb", "resnetb", "resnetb_strided", "resnetb", "resnetb",
               "resnetb_strided", "resnetb", "resnetb", "nearest_upsample", "unary", "nearest_upsample", "unary",
               "nearest_upsample", "last_unary"],
}
ARCHITECTURES["kitti"] = ARCHITECTURES["indoor"]


def indoor_config(**over):
    """configs/test/indoor.yaml:28-45 (model section)."""
    c = Config(num_layers=4, in_points_dim=3, first_feats_dim=256, final_feats_dim=32, first_subsampling_dl=0.025,
               in_feats_dim=1, conv_radius=2.5, deform_radius=5.0, num_kernel_points=15, KP_extent=2.0, KP_influence="linear",
               aggregation_mode="sum", fixed_kernel_points="center", use_batch_norm=True, batch_norm_momentum=0.02,
               deformable=False, modulated=False, architecture=ARCHITECTURES["indoor"],
               gnn_feats_dim=512, dgcnn_k=10, num_head=4, nets=["self", "cross", "self"])      # configs/test/indoor.yaml:31,48-50
    c.update(over)
    return c


def kitti_config(**over):
    """configs/test/kitti.yaml:10-27"""
    return indoor_config(**{**dict(first_subsampling_dl=0.3, conv_radius=4.25, gnn_feats_dim=256), **over})


class KPEncoder(nn.Module):
    """The encoder half of KPFCNN (models/architectures.py:62-100 construction, :520-524 loop).
    ``encoder_blocks`` has the reference's module names, so ``KPFCNN.state_dict()`` entries
    ``encoder_blocks.*`` load with ``load_state_dict(strict=False)`` / :meth:`load_reference`."""

    def __init__(self, config):
        super().__init__()
        layer, r = 0, config.first_subsampling_dl * config.conv_radius
        in_dim, out_dim = config.in_feats_dim, config.first_feats_dim
        self.encoder_blocks = nn.ModuleList()
        self.encoder_skips, self.encoder_skip_dims = [], []
        for block_i, block in enumerate(config.architecture):
            if any(t in block for t in ("pool", "strided", "upsample", "global")):
                self.encoder_skips.append(block_i)
                self.encoder_skip_dims.append(in_dim)
            if "upsample" in block:
                break
            self.encoder_blocks.append(block_decider(block, r, in_dim, out_dim, layer, config))
            in_dim = out_dim // 2 if "simple" in block else out_dim
            if "pool" in block or "strided" in block:
                layer += 1
                r *= 2
                out_dim *= 2
        self.out_dim = in_dim

    def load_reference(self, state_dict, prefix="encoder_blocks."):
        own = self.state_dict()
        for k in own:
            src = prefix + k[len("encoder_blocks."):]
            if src not in state_dict:
                raise KeyError(f"missing {src} in reference state_dict")
            own[k].copy_(torch.as_tensor(state_dict[src]))
        return self

    @torch.no_grad()
    def forward(self, x, batch, return_skips=False):
        skips = []
        for block_i, block_op in enumerate(self.encoder_blocks):
            if block_i in self.encoder_skips:
                skips.append(x)
            x = block_op(x, batch)
        return (x, skips) if return_skips else x


class KPDecoder(nn.Module):
    """The decoder tail of KPFCNN (construction models/architectures.py:114-153, loop :567-582):
    nearest_upsample -> concat skip -> unary ... -> last_unary, then the descriptor head.  Module names equal the
    reference's (``decoder_blocks.<i>.mlp.weight``), so ``KPFCNN.state_dict()`` entries load unchanged.  The
    bottleneck GNN between encoder and decoder stays with the reference (out of the hot-path scope)."""

    def __init__(self, config, encoder, gnn_feats_dim):
        super().__init__()
        arch = config.architecture
        start_i = next(i for i, b in enumerate(arch) if "upsample" in b)
        n_strided = sum(1 for b in arch[:start_i] if "pool" in b or "strided" in b)
        layer = n_strided
        r = config.first_subsampling_dl * config.conv_radius * (2 ** n_strided)
        in_dim, out_dim = encoder.out_dim, gnn_feats_dim + 2
        self.final_feats_dim = config.final_feats_dim
        self.decoder_blocks = nn.ModuleList()
        self.decoder_concats, self.block_layers = [], []
        for block_i, block in enumerate(arch[start_i:]):
            if block_i > 0 and "upsample" in arch[start_i + block_i - 1]:
                in_dim += encoder.encoder_skip_dims[layer]
                self.decoder_concats.append(block_i)
            self.decoder_blocks.append(block_decider(block, r, in_dim, out_dim, layer, config))
            self.block_layers.append(layer)
            in_dim = out_dim
            if "upsample" in block:
                layer -= 1
                r *= 0.5
                out_dim = out_dim // 2

    def load_reference(self, state_dict, prefix="decoder_blocks."):
        own = self.state_dict()
        for k in own:
            src = prefix + k[len("decoder_blocks."):]
            if src not in state_dict:
                raise KeyError(f"missing {src} in reference state_dict")
            own[k].copy_(torch.as_tensor(state_dict[src]))
        return self

    @torch.no_grad()
    def forward(self, x, skips, batch):
        """x [N_coarse, gnn_feats_dim + 2] (scores_c_raw, scores_saliency, feats_gnn_raw), skips = the encoder's skip list.
        -> (feats_f [N0, final_feats_dim], scores_overlap [N0], scores_saliency [N0])"""
        skips = list(skips)
        for block_i, block_op in enumerate(self.decoder_blocks):
            if block_i in self.decoder_concats:
                x = torch.cat([x, skips.pop()], dim=1)
            if isinstance(block_op, UnaryBlock):
                x = block_op(x, batch, segments=_segments(batch, self.block_layers[block_i]))
            else:
                x = block_op(x, batch)
        return ops.descriptor_head(x, self.final_feats_dim)
