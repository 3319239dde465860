"""The hot path end to end, as one object: stacked fragment pairs in -> encoder features out.

    subsample x3 + radius search x10  (collate_fn_descriptor, datasets/dataloader.py:239-359)
    -> 11 encoder blocks               (KPFCNN.forward encoder loop, models/architectures.py:520-524)

``run_device`` keeps everything in HBM; ``run_host`` is the user-facing call with HOST buffers
(pinned host -> device copy of the raw points, device -> host copy of the per-point features of
the coarsest level), which is what bench.py times for its ``e2e`` figure.
"""
import numpy as np
import torch

from . import blocks, dataloader

# neighbourhood limits frozen with the reference's calibration rule (datasets/dataloader.py:402-434,
# keep_ratio 0.8) run through the reference C++ core on the synthetic families of synthetic.py
CALIBRATED_LIMITS = {
    "3dmatch_synthetic": [34, 39, 39, 38],
    "3dlomatch_synthetic": [33, 40, 42, 40],
    "3dmatch_lattice_synthetic": [34, 39, 41, 39],
    "kitti_synthetic": [102, 102, 99, 91],
    "3dmatch_demo_pair": [38, 36, 36, 38],      # SURVEY.md section 0.7 (widely used 3DMatch setting)
}


def init_kernel_points(net, seed=0):
    """Deterministic stand-in for kernels/kernel_points.py:388-470 when no checkpoint is loaded:
    centre point + 14 points near the sphere of radius 0.66*conv radius (random-init weights of the
    named architecture; a real run loads them from the reference state_dict)."""
    g = torch.Generator().manual_seed(seed)
    for m in net.modules():
        if isinstance(m, blocks.KPConv):
            d = torch.randn(m.K, 3, generator=g)
            d = d / d.norm(dim=1, keepdim=True) * 0.66
            d[0] = 0
            m.set_kernel_points((d + 0.01 * torch.randn(m.K, 3, generator=g)) * m.radius)
    return net


class FeaturePath:
    def __init__(self, config=None, limits=None, device="cuda", state_dict=None, seed=0):
        self.config = config if config is not None else blocks.indoor_config()
        self.limits = list(limits if limits is not None else CALIBRATED_LIMITS["3dmatch_synthetic"])
        self.device = torch.device(device)
        torch.manual_seed(seed)
        self.encoder = blocks.KPEncoder(self.config)
        if state_dict is not None:
            self.encoder.load_reference(state_dict)
        else:
            init_kernel_points(self.encoder, seed)
        self.encoder.to(self.device).eval()

    @torch.no_grad()
    def run_device(self, points, lengths, features=None, views_per_cloud=None):
        """points [N,3] f32 cuda, lengths [2P] i32 cuda -> (features of the coarsest level [N3, C], batch dict), ready on the
        caller's current stream.  views_per_cloud (colour path, in_feats_dim = C2d + 1): per cloud its RGB-D views in write
        order (see projection.unproject_features_batch); the 2D features are un-projected onto the points and form the input
        rows."""
        return self.submit_device(points, lengths, features, views_per_cloud).result()

    def _streams(self):
        if not hasattr(self, "_s_pyr"):
            self._s_pyr = torch.cuda.Stream(device=self.device)
            self._s_enc = torch.cuda.Stream(device=self.device)
        return self._s_pyr, self._s_enc

    @torch.no_grad()
    def submit_device(self, points, lengths, features=None, views_per_cloud=None):
        """Asynchronous form of :meth:`run_device`: the pyramid (subsampling + searches, whose size read-backs block the host)
        runs on one stream and the encoder on another, so the pyramid of the NEXT submission -- and its host waits -- overlaps
        the encoder of this one instead of draining the GPU four times per step.  ``handle.result()`` makes the caller's
        current stream wait for the encoder and returns (features, batch)."""
        cur = torch.cuda.current_stream(self.device)
        sp, se = self._streams()
        ev_in = torch.cuda.Event()
        ev_in.record(cur)
        sp.wait_event(ev_in)
        with torch.cuda.stream(sp):
            if views_per_cloud is not None:
                from . import projection
                features = projection.unproject_features_batch(points, lengths, views_per_cloud)
            batch = dataloader.build_pyramid(points, lengths, self.config, self.limits, device=self.device)
            if features is None:
                features = torch.ones((points.shape[0], self.config.in_feats_dim), dtype=torch.float32, device=self.device)
            ev_p = torch.cuda.Event()
            ev_p.record(sp)
        se.wait_event(ev_p)
        with torch.cuda.stream(se):
            y = self.encoder(features, batch)
            ev_e = torch.cuda.Event()
            ev_e.record(se)
        # tensors of one stream's allocator pool that another stream reads
        for t in [points, lengths, features] + [x for k in ("points", "neighbors", "pools", "upsamples", "stack_lengths", "pair_segments")
                                                for x in batch[k]]:
            if torch.is_tensor(t) and t.is_cuda:
                t.record_stream(sp)
                t.record_stream(se)
        return _DeviceResult(y, batch, ev_e, self.device)

    @torch.no_grad()
    def run_host(self, points_host, lengths_host, out_host=None, views_per_cloud=None):
        """HOST buffers in and out.  points_host [N,3] f32 (pinned for async copies), lengths_host [2P] i32.
        Returns (features_host [N3,C], lengths of the coarsest level [2P] on the host)."""
        h = self.submit_host(points_host, lengths_host, out_host, views_per_cloud)
        return h.result()

    @torch.no_grad()
    def submit_host(self, points_host, lengths_host, out_host=None, views_per_cloud=None):
        """Pipelined form of :meth:`run_host`: the device->host copy of the result runs on a separate copy
        stream, so it overlaps the next submission's compute.  Returns a handle; ``handle.result()`` waits for
        this submission only.  ``out_host`` (pinned, reused by the caller once result() returned) avoids a
        pageable copy."""
        if not hasattr(self, "_copy_stream"):
            self._copy_stream = torch.cuda.Stream(device=self.device)
            self._pinned_small, self._slot = {}, 0
        pts = points_host.to(self.device, non_blocking=True)
        lens = lengths_host.to(self.device, non_blocking=True)
        h = self.submit_device(pts, lens, views_per_cloud=views_per_cloud)
        y, batch, ready = h.y, h.batch, h.ready
        coarse_dev = batch["stack_lengths"][-1]
        with torch.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(ready)
            if out_host is not None and out_host.shape[0] >= y.shape[0]:
                out = out_host[:y.shape[0]]
            else:
                out = torch.empty(y.shape, dtype=y.dtype, pin_memory=True)
            # in 4 MB pieces: the pyramid's small size read-backs of the NEXT submission share the D2H copy
            # engine and would otherwise queue behind one monolithic transfer
            step = max(1, (4 << 20) // max(1, y.shape[1] * 4))
            for r in range(0, y.shape[0], step):
                out[r:r + step].copy_(y[r:r + step], non_blocking=True)
            self._slot = (self._slot + 1) % 4    # small pinned staging buffers are cached (pinned allocation is slow); 4 in flight
            key = (self._slot, tuple(coarse_dev.shape))
            if key not in self._pinned_small:
                self._pinned_small[key] = torch.empty(coarse_dev.shape, dtype=coarse_dev.dtype, pin_memory=True)
            coarse = self._pinned_small[key]
            coarse.copy_(coarse_dev, non_blocking=True)
            y.record_stream(self._copy_stream)
            coarse_dev.record_stream(self._copy_stream)
            done = torch.cuda.Event()
            done.record(self._copy_stream)
        return _HostResult(out, coarse, done)


class _DeviceResult:
    def __init__(self, y, batch, ready, device):
        self.y, self.batch, self.ready, self._device = y, batch, ready, device

    def result(self):
        cur = torch.cuda.current_stream(self._device)
        cur.wait_event(self.ready)
        self.y.record_stream(cur)
        return self.y, self.batch


class _HostResult:
    def __init__(self, out, coarse, done):
        self._out, self._coarse, self._done = out, coarse, done

    def result(self):
        self._done.synchronize()
        return self._out, self._coarse


def stack_pairs(pairs):
    """[(src, tgt), ...] -> (points [N,3] f32, lengths [2P] i32) NumPy, in the reference's src,tgt order."""
    pts = np.concatenate([np.concatenate([s, t]) for s, t in pairs]).astype(np.float32)
    lens = np.array([len(c) for p in pairs for c in p], np.int32)
    return pts, lens


def per_pair_features(path, pairs):
    """compute() for sharding.run_sharded: runs one stacked batch and splits the coarsest-level
    features back into per-pair tensors."""
    pts, lens = stack_pairs(pairs)
    dev = path.device
    y, batch = path.run_device(torch.from_numpy(pts).to(dev), torch.from_numpy(lens).to(dev))
    seg = batch["pair_segments"][-1].cpu().tolist()
    return [y[seg[k]:seg[k + 1]] for k in range(len(pairs))]
