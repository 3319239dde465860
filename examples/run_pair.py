#!/usr/bin/env python
"""One fragment pair through the whole B200 path, the way the reference's demo / tester drives it
(scripts/demo.py, lib/tester.py:52-102): pyramid (grid subsampling + radius searches) -> KPFCNN (encoder, overlap-attention
bottleneck, decoder) -> per-point descriptors and scores -> mutual descriptor matching -> the tester's per-pair .pth file.

    python examples/run_pair.py [--weights reference_checkpoint.pth] [--out pair0.pth]

Without --weights the network is randomly initialised (there are no checkpoints in this repository); with a PCR-CG /
Predator checkpoint its `state_dict` loads unchanged (same parameter names).  Input: tests/golden/demo_pair_f32.npz, the
reference's own demo fragments (assets/cloud_bin_21.pth / cloud_bin_34.pth as float32)."""
import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pcrcg_b200 import architectures, blocks, dataloader, matching, pipeline  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--weights", default=None, help="reference checkpoint (.pth with a 'state_dict' entry)")
    ap.add_argument("--out", default=None, help="write the tester's per-pair file here")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    z = np.load(os.path.join(ROOT, "tests", "golden", "demo_pair_f32.npz"))
    src, tgt = z["src"], z["tgt"]
    cfg = blocks.indoor_config()                                   # configs/test/indoor.yaml (in_feats_dim 1: geometry only)
    net = architectures.KPFCNN(cfg)
    if a.weights:
        sd = torch.load(a.weights, map_location="cpu", weights_only=False)
        net.load_state_dict(sd.get("state_dict", sd), strict=True)
    else:
        pipeline.init_kernel_points(net, 0)
    net.to(dev).eval()

    pts = np.concatenate([src, tgt]).astype(np.float32)
    lens = np.array([len(src), len(tgt)], np.int32)
    limits = pipeline.CALIBRATED_LIMITS["3dmatch_demo_pair"]
    for it in range(2):                                            # second pass = warm timings
        torch.cuda.synchronize(); t0 = time.perf_counter()
        batch = dataloader.build_pyramid(pts, lens, cfg, limits, device=dev)
        torch.cuda.synchronize(); t1 = time.perf_counter()
        batch["features"] = torch.ones(len(pts), cfg.in_feats_dim, device=dev)
        res = net(batch)
        torch.cuda.synchronize(); t2 = time.perf_counter()
    feats, ov, sal = res["feats_f"], res["scores_overlap"], res["scores_saliency"]
    n_src = len(src)
    rows, cols = matching.mutual_matches(feats[:n_src], feats[n_src:])
    torch.cuda.synchronize(); t3 = time.perf_counter()
    print(f"levels {[int(p.shape[0]) for p in batch['points']]}  pyramid {1e3 * (t1 - t0):.2f} ms  network {1e3 * (t2 - t1):.2f} ms  "
          f"mutual matching {1e3 * (t3 - t2):.2f} ms")
    print(f"descriptors {tuple(feats.shape)}  |f| = {float(feats.norm(dim=1).mean()):.4f}  overlap mean {float(ov.mean()):.3f}  "
          f"saliency mean {float(sal.mean()):.3f}  mutual matches {int(rows.numel())}")
    if a.out:
        matching.save_pair(a.out, batch["points"][0], feats, ov, sal, n_src, torch.eye(3), torch.zeros(3, 1))
        print("wrote", a.out, sorted(matching.load_pair(a.out).keys()))


if __name__ == "__main__":
    main()
