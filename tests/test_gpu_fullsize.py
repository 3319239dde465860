"""GPU, full size (config 1 of BASELINE.json: the reference's demo pair cloud_bin_21 / cloud_bin_34):
the whole pyramid must reproduce the digests of the REFERENCE's own run (tests/golden/preprocess_ref.npz,
made by tests/golden/make_golden.py with limits 38/36/36/38), and the 11-block encoder with
first_feats_dim = 256 must stay within 1e-3 (normwise) of the PyTorch-fp32 CPU restatement fed the same
index lists and the same weights."""
import hashlib
import os

import numpy as np
import pytest
import torch

from oracle import blocks_port as bp
from pcrcg_b200 import blocks, dataloader, pipeline

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
G = os.path.join(os.path.dirname(__file__), "golden")


def sha(t):
    a = t.detach().cpu().numpy() if torch.is_tensor(t) else np.asarray(t)
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def demo_batch(demo_pair):
    pre = np.load(os.path.join(G, "preprocess_ref.npz"))
    pts = np.concatenate(demo_pair)
    lens = np.array([len(demo_pair[0]), len(demo_pair[1])], np.int32)
    cfg = blocks.indoor_config()
    b = dataloader.build_pyramid(pts, lens, cfg, pre["demo_limits"].tolist(), device=DEV)
    return pre, cfg, b


def test_demo_pair_pyramid_equals_reference_digests(demo_batch):
    pre, cfg, b = demo_batch
    assert [int(p.shape[0]) for p in b["points"]] == pre["demo_level_sizes"].tolist() == [39939, 9932, 2612, 758]
    assert np.array_equal(torch.stack(b["stack_lengths"]).cpu().numpy(), pre["demo_stack_lengths"])
    for l in range(4):
        assert sha(b["points"][l]) == str(pre[f"demo_points_sha_{l}"]), f"points level {l}"
        assert sha(b["neighbors"][l].contiguous()) == str(pre[f"demo_neighbors_sha_{l}"]), f"neighbors level {l}"
        if l < 3:
            assert sha(b["pools"][l].contiguous()) == str(pre[f"demo_pools_sha_{l}"]), f"pools level {l}"
            assert sha(b["upsamples"][l].contiguous()) == str(pre[f"demo_upsamples_sha_{l}"]), f"upsamples level {l}"


def test_demo_pair_encoder_256_vs_cpu_port(demo_batch):
    pre, cfg, b = demo_batch
    torch.manual_seed(3)
    enc = blocks.KPEncoder(cfg)
    pipeline.init_kernel_points(enc, 3)
    sd = {k: v.clone() for k, v in enc.state_dict().items()}
    enc.to(DEV)
    n0 = b["points"][0].shape[0]
    y = enc(torch.ones(n0, 1, device=DEV), b).cpu()
    cpu_batch = {k: [t.cpu().contiguous() for t in b[k]] for k in ("points", "neighbors", "pools", "upsamples")}
    desc = bp.encoder_blocks_from_state_dict(sd, prefix="encoder_blocks.")
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    with torch.no_grad():
        ref, _ = bp.encoder(torch.ones(n0, 1), cpu_batch, desc)
    assert y.shape == ref.shape == (758, 2048)
    err = float((y - ref).abs().max() / ref.abs().max())
    assert err < 1e-3, err


def test_stacked_demo_pairs_roundtrip(demo_pair):
    """Size-independent property at batch size: 6 stacked copies of the demo pair give 6 identical pyramids."""
    cfg = blocks.indoor_config()
    pts = np.concatenate([np.concatenate(demo_pair)] * 6)
    lens = np.array([len(demo_pair[0]), len(demo_pair[1])] * 6, np.int32)
    b = dataloader.build_pyramid(pts, lens, cfg, [38, 36, 36, 38], device=DEV)
    for l in range(4):
        n = b["points"][l].shape[0] // 6
        p = b["points"][l].view(6, n, 3)
        assert torch.equal(p[0].expand_as(p), p)
        nb = b["neighbors"][l].contiguous().view(6, n, -1).long()
        off = (torch.arange(6, device=DEV) * n).view(6, 1, 1)
        shadow = nb >= 6 * n
        rel = torch.where(shadow, torch.full_like(nb, -1), nb - off)
        assert torch.equal(rel[0].expand_as(rel), rel)


def test_kpconv_support_permutation_invariance_full_size(demo_batch):
    """Size-independent property at full size (39 939 query points x 38 neighbours, 64 -> 64 channels): relabelling the
    support rows (features, coordinates and the index lists remapped consistently) must not change ANY output bit -- the
    neighbour order inside each list, hence every summation order, is unchanged; only the gather addresses move.  Run on both
    aggregation kernels (pipelined bf16 planes, and the fp32-input tensor-core kernel)."""
    from pcrcg_b200 import ops
    _, cfg, b = demo_batch
    pts, idx = b["points"][0], b["neighbors"][0]
    n = pts.shape[0]
    g = torch.Generator().manual_seed(11)
    raw = torch.randn(n, 64, generator=g).to(DEV)
    w = (torch.randn(15, 64, 64, generator=g) / 31.0).to(DEV)
    kp = (torch.randn(15, 3, generator=g) * 0.03).to(DEV)
    perm = torch.randperm(n, generator=g).to(DEV)
    inv = torch.empty_like(perm)
    inv[perm] = torch.arange(n, device=DEV)
    inv_pad = torch.cat([inv, torch.tensor([n], device=DEV)])                    # the shadow index stays the shadow index
    idx_p = inv_pad[idx.long()].to(torch.int32)
    for planes in (True, False):
        x = ops.instance_norm_act(raw, None, 0.1, emit_split=planes, emit_rowpos=planes)
        xp = x[perm].contiguous()                       # the SAME feature values, rows relabelled (planes and row flags too)
        if planes:
            hi, lo, ld = ops.attached(x, "_pcrcg_split")
            ops._attach(xp, "_pcrcg_split", (hi[perm].contiguous(), lo[perm].contiguous(), ld))
            ops._attach(xp, "_pcrcg_rowpos", ops.attached(x, "_pcrcg_rowpos")[perm].contiguous())
        out = ops.kpconv_forward(pts, pts, idx, x, kp, w, 0.05)
        out_p = ops.kpconv_forward(pts, pts[perm], idx_p, xp, kp, w, 0.05)
        assert torch.equal(out, out_p), f"planes={planes}"
    assert float(out.abs().max()) > 0


def test_demo_pair_whole_network_vs_cpu_ports(demo_batch):
    """config 1 at full size through the WHOLE descriptor network (indoor.yaml dims: encoder 256 -> 2048, GNN 512, 4 heads,
    k = 10, decoder -> 32 + 2): final descriptors and scores vs the PyTorch-fp32 CPU restatements (oracle/blocks_port.py,
    oracle/gcn_port.py; each pinned to the reference's own outputs) with the same weights and the same index lists."""
    from oracle import gcn_port as gp
    from pcrcg_b200 import architectures
    pre, cfg, b = demo_batch
    torch.manual_seed(5)
    net = architectures.KPFCNN(cfg)
    for m in net.encoder_blocks.modules():
        if isinstance(m, blocks.KPConv):
            m.set_kernel_points(torch.randn(15, 3) * 0.4 * m.radius)
    with torch.no_grad():
        net.epsilon.fill_(-2.0)
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    net.to(DEV)
    n0 = b["points"][0].shape[0]
    fb = dict(b)
    fb["features"] = torch.ones(n0, 1, device=DEV)
    res = {k: v.cpu() for k, v in net(fb).items()}
    assert res["feats_f"].shape == (n0, 32)
    # CPU ports, stage by stage
    cpu_batch = {k: [t.cpu().contiguous() for t in b[k]] for k in ("points", "neighbors", "pools", "upsamples")}
    desc = bp.encoder_blocks_from_state_dict(sd, prefix="encoder_blocks.")
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    with torch.no_grad():
        x, outs = bp.encoder(torch.ones(n0, 1), cpu_batch, desc)
        xb = gp.bottleneck(x, cpu_batch["points"][3], b["stack_lengths"][3].cpu().numpy(), sd, cfg.nets, cfg.num_head, cfg.dgcnn_k)
        Ws = [sd["decoder_blocks.1.mlp.weight"], sd["decoder_blocks.3.mlp.weight"], sd["decoder_blocks.5.mlp.weight"]]
        feats, so, ss, _ = bp.decoder(xb, [outs[1], outs[4], outs[7]], cpu_batch, Ws, 32)
    err = lambda a, r: float((a - r).abs().max() / r.abs().max().clamp_min(1e-30))
    assert err(res["feats_f"], feats) < 1e-3, err(res["feats_f"], feats)
    assert err(res["scores_overlap"], so) < 1e-3 and err(res["scores_saliency"], ss) < 1e-3
