"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): PyTorch-CPU restatement of the bottleneck of
PCR-CG's KPFCNN -- the overlap-attention GNN (models/gcn.py) and the projections / saliency scores
around it (models/architectures.py:528-565).  Row-major [N, C] tensors (the reference keeps
[1, C, N]); every function cites the reference lines it follows.  Pinned against the reference itself
by tests/test_oracle_pinning.py through tests/golden/gnn_ref.npz (made by tests/golden/make_golden.py).
"""
import torch
import torch.nn.functional as F


def square_distance(src, dst):
    """models/gcn.py:16-35 (normalised=False): -2 x.y + |x|^2 + |y|^2, clamped at 1e-12."""
    dist = -2 * torch.matmul(src, dst.t())
    dist += torch.sum(src ** 2, dim=-1)[:, None]
    dist += torch.sum(dst ** 2, dim=-1)[None, :]
    return torch.clamp(dist, min=1e-12)


def knn_lists(coords, lens, k):
    """models/gcn.py:48-51 per cloud: the k+1 smallest distances, first one (the query itself) dropped.
    -> int64 [N, k] of GLOBAL row indices."""
    out, off = [], 0
    for n in lens:
        c = coords[off:off + int(n)]
        idx = square_distance(c, c).topk(k=k + 1, dim=-1, largest=False, sorted=True)[1][:, 1:]
        out.append(idx + off)
        off += int(n)
    return torch.cat(out)


def _inorm(x, eps=1e-5):
    """nn.InstanceNorm1d / 2d defaults: no affine, biased variance over every non-channel position (rows here)."""
    mu = x.mean(0, keepdim=True)
    var = x.var(0, unbiased=False, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps)


def _edge_conv(feats, idx, w):
    """get_graph_feature + 1x1 conv + InstanceNorm2d + LeakyReLU(0.2) + max over the k edges
    (models/gcn.py:54-66, 125-127).  feats [n, C], idx [n, k] local rows, w [Cout, 2C]."""
    n, k = idx.shape
    centre = feats[:, None, :].expand(n, k, feats.shape[1])
    edge = torch.cat([centre, feats[idx] - centre], dim=-1)                  # [n, k, 2C]
    y = edge.reshape(n * k, -1) @ w.t()                                       # conv over every (node, edge)
    y = F.leaky_relu(_inorm(y), 0.2).reshape(n, k, -1)
    return y.max(dim=1)[0]


def self_attention(coords, feats, sd, prefix, k):
    """SelfAttention.forward (models/gcn.py:113-137) on ONE cloud."""
    idx = knn_lists(coords, [len(coords)], k)
    w1 = sd[prefix + "conv1.weight"].reshape(sd[prefix + "conv1.weight"].shape[0], -1)
    w2 = sd[prefix + "conv2.weight"].reshape(sd[prefix + "conv2.weight"].shape[0], -1)
    w3 = sd[prefix + "conv3.weight"].reshape(sd[prefix + "conv3.weight"].shape[0], -1)
    x1 = _edge_conv(feats, idx, w1)
    x2 = _edge_conv(x1, idx, w2)
    x3 = torch.cat([feats, x1, x2], dim=1) @ w3.t()
    return F.leaky_relu(_inorm(x3), 0.2)


def _conv1d(x, sd, name):
    w = sd[name + ".weight"]
    return x @ w.reshape(w.shape[0], -1).t() + sd[name + ".bias"]


def attentional_propagation(x, source, sd, prefix, num_heads):
    """AttentionalPropagation.forward (models/gcn.py:183-185) with MultiHeadedAttention (:160-172) and
    attention (:153-157).  Channel c of a projection is (d = c // heads, h = c % heads) (the .view of :168)."""
    q = _conv1d(x, sd, prefix + "attn.proj.0")
    kk = _conv1d(source, sd, prefix + "attn.proj.1")
    v = _conv1d(source, sd, prefix + "attn.proj.2")
    C = q.shape[1]
    dim = C // num_heads
    qh, kh, vh = (t.reshape(t.shape[0], dim, num_heads) for t in (q, kk, v))       # [n, d, h]
    scores = torch.einsum("ndh,mdh->hnm", qh, kh) / dim ** 0.5
    prob = torch.softmax(scores, dim=-1)
    msg = torch.einsum("hnm,mdh->ndh", prob, vh).reshape(-1, C)
    msg = _conv1d(msg, sd, prefix + "attn.merge")
    y = _conv1d(torch.cat([x, msg], dim=1), sd, prefix + "mlp.0")
    y = torch.relu(_inorm(y))
    return _conv1d(y, sd, prefix + "mlp.3")


def gcn(coords, lens, feats, sd, names, num_heads, k, prefix=""):
    """GCN.forward (models/gcn.py:207-217).  coords [N,3], feats [N,C] stacked (src, tgt), lens = (n_src, n_tgt)."""
    n0 = int(lens[0])
    c0, c1, d0, d1 = coords[:n0], coords[n0:], feats[:n0], feats[n0:]
    for i, name in enumerate(names):
        p = f"{prefix}layers.{i}."
        if name == "cross":
            d0 = d0 + attentional_propagation(d0, d1, sd, p, num_heads)
            d1 = d1 + attentional_propagation(d1, d0, sd, p, num_heads)
        else:
            d0 = self_attention(c0, d0, sd, p, k)
            d1 = self_attention(c1, d1, sd, p, k)
    return torch.cat([d0, d1])


def bottleneck(x, coords_c, lens_c, sd, names, num_heads, k):
    """models/architectures.py:528-565: bottle -> GNN -> proj_gnn / proj_score -> saliency scores.
    x [Nc, C_enc] encoder output.  -> decoder input [Nc, 2 + gnn_feats_dim] = (scores_c_raw, scores_saliency, feats_gnn_raw)."""
    n0 = int(lens_c[0])
    f = _conv1d(x, sd, "bottle")
    f = gcn(coords_c, lens_c, f, sd, names, num_heads, k, prefix="gnn.")
    f = _conv1d(f, sd, "proj_gnn")
    scores = _conv1d(f, sd, "proj_score")                                          # [Nc, 1]
    fn = F.normalize(f, p=2, dim=1)
    inner = fn[:n0] @ fn[n0:].t()
    temperature = torch.exp(sd["epsilon"]) + 0.03
    s1 = torch.softmax(inner / temperature, dim=1) @ scores[n0:]
    s2 = torch.softmax(inner.t() / temperature, dim=1) @ scores[:n0]
    return torch.cat([scores, torch.cat([s1, s2]), f], dim=1)


# ---- node labels of the collate (datasets/dataloader.py:91-198) --------------------------------------------------------
def point2node(nodes, points):
    """datasets/dataloader.py:91-106 (square_distance :70-90 is the same expanded form as models/gcn.py)."""
    return square_distance(points, nodes).topk(k=1, dim=-1, largest=False)[1].squeeze(-1)


def point2node_correspondences(src_nodes, src_points, tgt_nodes, tgt_points, corr):
    """datasets/dataloader.py:108-198: visible fraction per node and the point -> node assignment."""
    out = []
    for nodes, pts, col in ((src_nodes, src_points, 0), (tgt_nodes, tgt_points, 1)):
        idx = point2node(nodes, pts)
        visible = torch.zeros(pts.shape[0], dtype=torch.bool)
        visible[corr[:, col]] = True
        tot = torch.bincount(idx, minlength=nodes.shape[0]).float()
        vis = torch.bincount(idx[visible], minlength=nodes.shape[0]).float()
        out.append((vis / torch.where(tot > 0, tot, torch.ones_like(tot)), idx))
    return out[0][0], out[1][0], out[0][1], out[1][1]
