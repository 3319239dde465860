"""Overlap-attention GNN of the KPFCNN bottleneck with the reference's module API (``models/gcn.py``), computed by the
CUDA kernels of libpcrcg_b200.so (csrc/gnn.cu + the tcgen05 / CUDA-core contraction).  Forward only.

Module and PARAMETER names equal the reference's, so ``KPFCNN.state_dict()`` entries ``gnn.layers.<i>.*`` load unchanged:
``conv{1,2,3}.weight`` (SelfAttention, models/gcn.py:101-108), ``attn.proj.{0,1,2}.{weight,bias}``, ``attn.merge.*``,
``mlp.{0,3}.*`` (AttentionalPropagation, :160-181).

Native layout is row-major ``[N, C]`` with stacked clouds; :meth:`GCN.forward` also accepts the reference's
``[1, C, N]`` / ``[1, 3, N]`` tensors.  Differences, all deliberate:
  * the edge convolution ``W [f_n ; f_j - f_n]`` is evaluated as ``(Wa - Wb) f_n + Wb f_j`` (two node-level contractions
    instead of one over N*k edges) and the max over edges is taken before the (increasing) norm + LeakyReLU;
  * heads are made contiguous by permuting the projection rows / merge columns once (the reference interleaves them,
    channel = d * heads + h, models/gcn.py:168), so a head is a column slice and needs no copy;
  * the bias of ``mlp.0`` is not added: the InstanceNorm1d that follows removes any per-channel constant;
  * kNN ties are broken by index (the reference's ``topk`` tie order is unspecified).
"""
import torch
import torch.nn as nn
from torch.nn.parameter import Parameter

from . import ops


class _Conv(nn.Module):
    """parameter holder with nn.Conv1d / nn.Conv2d (kernel size 1) names and shapes"""

    def __init__(self, cin, cout, nd, bias):
        super().__init__()
        self.weight = Parameter(torch.empty((cout, cin) + (1,) * nd, dtype=torch.float32), requires_grad=False)
        nn.init.kaiming_uniform_(self.weight.view(cout, cin), a=5 ** 0.5)
        self.bias = Parameter(torch.zeros(cout, dtype=torch.float32), requires_grad=False) if bias else None
        if bias:
            nn.init.uniform_(self.bias, -1.0 / cin ** 0.5, 1.0 / cin ** 0.5)

    @property
    def w2d(self):
        return self.weight.view(self.weight.shape[0], self.weight.shape[1])


class _Cache:
    """derived weights, rebuilt when a source parameter changes"""

    def __init__(self):
        self.key, self.val = None, None

    def get(self, params, build):
        key = tuple((p.data_ptr(), p._version) for p in params)
        if key != self.key:
            self.key, self.val = key, build()
        return self.val


class SelfAttention(nn.Module):
    """models/gcn.py:98-137 (DGCNN-style graph features, k nearest neighbours in coordinate space)."""

    def __init__(self, feature_dim, k=10):
        super().__init__()
        self.conv1 = _Conv(feature_dim * 2, feature_dim, 2, False)
        self.conv2 = _Conv(feature_dim * 2, feature_dim * 2, 2, False)
        self.conv3 = _Conv(feature_dim * 4, feature_dim, 2, False)
        self.k, self.feature_dim = k, feature_dim
        self._cache = _Cache()

    def _split_weights(self):
        def build():
            out = []
            for conv in (self.conv1, self.conv2):
                w = conv.w2d
                c = w.shape[1] // 2
                out.append(torch.cat([w[:, :c] - w[:, c:], w[:, c:]], dim=0).contiguous())     # rows: (Wa - Wb | Wb)
            return out
        return self._cache.get([self.conv1.weight, self.conv2.weight], build)

    def forward_rows(self, coords, starts, feats, knn_idx=None):
        """coords [N,3], feats [N,C] of stacked clouds with row starts ``starts`` (statistics per cloud).  -> [N,C]"""
        C = self.feature_dim
        if knn_idx is None:
            knn_idx = ops.knn(coords, starts, self.k)
        w1, w2 = self._split_weights()
        seg = starts.to(feats.device)
        x1 = ops.edge_conv_max(ops.linear(feats, w1), C, knn_idx, seg)
        x2 = ops.edge_conv_max(ops.linear(x1, w2), 2 * C, knn_idx, seg)
        x3 = ops.linear(torch.cat([feats, x1, x2], dim=1), self.conv3.w2d, stat_segments=seg)
        return ops.instance_norm_act(x3, seg, 0.2)

    def forward(self, coords, features):
        """reference layout: coords [1,3,N], features [1,C,N] -> [1,C,N]"""
        n = features.shape[2]
        out = self.forward_rows(coords[0].t().contiguous(), ops.cloud_starts([n]), features[0].t().contiguous())
        return out.t().contiguous().unsqueeze(0)


class MultiHeadedAttention(nn.Module):
    """models/gcn.py:160-172"""

    def __init__(self, num_heads, d_model):
        super().__init__()
        assert d_model % num_heads == 0
        self.dim, self.num_heads = d_model // num_heads, num_heads
        self.merge = _Conv(d_model, d_model, 1, True)
        self.proj = nn.ModuleList([_Conv(d_model, d_model, 1, True) for _ in range(3)])
        self._cache = _Cache()

    def _head_major(self):
        def build():
            H, D = self.num_heads, self.dim
            # new channel h * D + d  <-  reference channel d * H + h
            perm = (torch.arange(D, device=self.merge.weight.device)[None, :] * H +
                    torch.arange(H, device=self.merge.weight.device)[:, None]).reshape(-1)
            pw = [(p.w2d[perm].contiguous(), p.bias[perm].contiguous()) for p in self.proj]
            return pw, self.merge.w2d[:, perm].contiguous()
        params = [self.merge.weight] + [p.weight for p in self.proj] + [p.bias for p in self.proj]
        return self._cache.get(params, build)

    def forward_rows(self, query, key, value):
        """query [n,C], key / value [m,C] -> [n,C]"""
        pw, merge_w = self._head_major()
        q, k, v = (ops.bias_act(ops.linear(x, w), b) for x, (w, b) in zip((query, key, value), pw))
        D = self.dim
        msg = torch.empty_like(q)
        for h in range(self.num_heads):
            sl = slice(h * D, (h + 1) * D)
            prob = ops.softmax_rows_(ops.gemm(q[:, sl], k[:, sl], True), 1.0 / D ** 0.5)       # models/gcn.py:153-156
            ops.gemm(prob, v[:, sl], False, out=msg[:, sl])
        return ops.bias_act(ops.linear(msg, merge_w), self.merge.bias)


class AttentionalPropagation(nn.Module):
    """models/gcn.py:175-185"""

    def __init__(self, feature_dim, num_heads):
        super().__init__()
        self.attn = MultiHeadedAttention(num_heads, feature_dim)
        # nn.Sequential(Conv1d, InstanceNorm1d, ReLU, Conv1d) of MLP([2C, 2C, C]) (models/gcn.py:140-150): same indices
        self.mlp = nn.ModuleList([_Conv(feature_dim * 2, feature_dim * 2, 1, True), nn.Identity(), nn.Identity(),
                                  _Conv(feature_dim * 2, feature_dim, 1, True)])
        nn.init.constant_(self.mlp[3].bias, 0.0)

    def forward_rows(self, x, source):
        message = self.attn.forward_rows(x, source, source)
        y = ops.linear(torch.cat([x, message], dim=1), self.mlp[0].w2d, stat_segments=True)
        y = ops.instance_norm_act(y, None, 0.0)                                  # InstanceNorm1d + ReLU
        return ops.bias_act(ops.linear(y, self.mlp[3].w2d), self.mlp[3].bias)


class GCN(nn.Module):
    """models/gcn.py:188-217: alternate self- and cross-attention."""

    def __init__(self, num_head, feature_dim, k, layer_names):
        super().__init__()
        self.layers = nn.ModuleList([AttentionalPropagation(feature_dim, num_head) if t == "cross" else SelfAttention(feature_dim, k)
                                     for t in layer_names])
        self.names = list(layer_names)
        self.k = k

    @torch.no_grad()
    def forward_rows(self, coords, lens, feats):
        """coords [N,3], feats [N,C]: stacked clouds (src_0, tgt_0, src_1, tgt_1, ...), lens [2P].  Self-attention layers run
        on all clouds at once (statistics per cloud), cross-attention per fragment pair.  -> [N,C]"""
        starts = ops.cloud_starts(lens)
        st = [int(s) for s in starts]
        knn_idx = ops.knn(coords, starts, self.k) if "self" in self.names else None
        for layer, name in zip(self.layers, self.names):
            if name == "self":
                feats = layer.forward_rows(coords, starts, feats, knn_idx)
                continue
            out = torch.empty_like(feats)
            for p in range(0, len(st) - 1, 2):
                d0, d1 = feats[st[p]:st[p + 1]], feats[st[p + 1]:st[p + 2]]
                d0 = ops.add_act(d0, layer.forward_rows(d0, d1), -1.0)           # desc0 = desc0 + layer(desc0, desc1)
                d1 = ops.add_act(d1, layer.forward_rows(d1, d0), -1.0)           # desc1 uses the UPDATED desc0 (:213-214)
                out[st[p]:st[p + 1]], out[st[p + 1]:st[p + 2]] = d0, d1
            feats = out
        return feats

    def forward(self, coords0, coords1, desc0, desc1):
        """reference layout: coords [1,3,N], desc [1,C,N] -> (desc0, desc1)"""
        n0, n1 = desc0.shape[2], desc1.shape[2]
        coords = torch.cat([coords0[0].t(), coords1[0].t()]).contiguous()
        feats = torch.cat([desc0[0].t(), desc1[0].t()]).contiguous()
        out = self.forward_rows(coords, [n0, n1], feats)
        return out[:n0].t().contiguous().unsqueeze(0), out[n0:].t().contiguous().unsqueeze(0)
