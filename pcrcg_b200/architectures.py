"""KPFCNN -- the whole descriptor network of PCR-CG (``models/architectures.py:33-610``) on the CUDA operators of
libpcrcg_b200.so: KPConv encoder -> bottleneck (``bottle`` -> overlap-attention GNN -> ``proj_gnn`` / ``proj_score`` ->
saliency scores) -> decoder -> descriptor head.  Forward only.

Parameter names equal the reference's (``encoder_blocks.*``, ``bottle.*``, ``gnn.layers.*``, ``proj_gnn.*``,
``proj_score.*``, ``epsilon``, ``decoder_blocks.*``): ``load_state_dict(reference_kpfcnn.state_dict())`` works unchanged.

Scope notes: the 2D image branch (``image_feature``; models/architectures.py:217-370) is fed through
``pcrcg_b200.projection`` by the caller, which hands this module ready ``batch['features']`` (N x 129); the
``node_overlap`` / ``quaternion`` heads (training-only side outputs, :545-552, :584-603) are not built.
"""
import torch
import torch.nn as nn
from torch.nn.parameter import Parameter

from . import ops
from .blocks import KPDecoder, KPEncoder
from .gcn import GCN, _Conv


class KPFCNN(nn.Module):
    def __init__(self, config):
        super().__init__()
        if dict.get(config, "node_overlap", False) or dict.get(config, "quaternion", False):
            raise NotImplementedError("pcrcg_b200.KPFCNN: the node_overlap / quaternion training heads are not built")
        enc = KPEncoder(config)
        self.encoder_blocks = enc.encoder_blocks
        self.encoder_skips, self.encoder_skip_dims = enc.encoder_skips, enc.encoder_skip_dims
        object.__setattr__(self, "_enc", enc)                    # not a registered submodule: names stay the reference's
        g = config.gnn_feats_dim
        self.bottle = _Conv(enc.out_dim, g, 1, True)             # models/architectures.py:106
        self.gnn = GCN(config.num_head, g, config.dgcnn_k, config.nets)
        self.proj_gnn = _Conv(g, g, 1, True)
        self.proj_score = _Conv(g, 1, 1, True)
        self.epsilon = Parameter(torch.tensor(-5.0), requires_grad=False)          # :55
        dec = KPDecoder(config, enc, g)
        self.decoder_blocks = dec.decoder_blocks
        object.__setattr__(self, "_dec", dec)
        self.final_feats_dim = config.final_feats_dim

    @torch.no_grad()
    def bottleneck(self, x, coords_c, lens_c):
        """models/architectures.py:528-565.  x [Nc, C_enc], coords_c [Nc,3], lens_c [2P] (src_0, tgt_0, ...).
        -> decoder input [Nc, 2 + gnn_feats_dim] = (scores_c_raw, scores_saliency, feats_gnn_raw)"""
        f = ops.bias_act(ops.linear(x, self.bottle.w2d), self.bottle.bias)
        f = self.gnn.forward_rows(coords_c, lens_c, f)
        f = ops.bias_act(ops.linear(f, self.proj_gnn.w2d), self.proj_gnn.bias)
        scores = ops.bias_act(ops.linear(f, self.proj_score.w2d), self.proj_score.bias)          # [Nc, 1]
        fn = ops.l2_normalize(f)
        inv_t = 1.0 / (float(torch.exp(self.epsilon)) + 0.03)
        sal = torch.empty_like(scores)
        st = [int(s) for s in ops.cloud_starts(lens_c)]
        for p in range(0, len(st) - 1, 2):
            a, b, c = st[p], st[p + 1], st[p + 2]
            w1 = ops.softmax_rows_(ops.gemm(fn[a:b], fn[b:c], True), inv_t)          # softmax(inner / T, dim=1)
            ops.gemm(w1, scores[b:c], False, out=sal[a:b])
            w2 = ops.softmax_rows_(ops.gemm(fn[b:c], fn[a:b], True), inv_t)          # softmax(inner^T / T, dim=1)
            ops.gemm(w2, scores[a:b], False, out=sal[b:c])
        return torch.cat([scores, sal, f], dim=1)

    @torch.no_grad()
    def forward(self, batch, backbone2d=None):
        """batch: the collate dict (``features``, ``points``, ``neighbors``, ``pools``, ``upsamples``, ``stack_lengths``
        [+ ``pair_segments`` for stacked pairs]) -> {'feats_f', 'scores_overlap', 'scores_saliency'} (:605-609)"""
        x = batch["features"]
        x, skips = self._enc(x, batch, return_skips=True)
        lens_c = batch["stack_lengths"][-1]
        xb = self.bottleneck(x, batch["points"][-1], lens_c)
        feats_f, scores_overlap, scores_saliency = self._dec(xb, skips, batch)
        return {"feats_f": feats_f, "scores_overlap": scores_overlap, "scores_saliency": scores_saliency}
