"""oracle.checks -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Parity of a STACKED GPU pyramid (many fragment pairs in one launch, pcrcg_b200.dataloader.build_pyramid) against the
reference run pair by pair on the CPU, the way datasets/dataloader.py:203-400 does it (one pair per call):

  * cpu_pyramid()   the reference's pyramid of ONE pair: subsample_batch + batch_query of the unmodified reference C++
                    (oracle/_ref) when it is built, else of the C port; neighbour rows in the canonical (d2, index) tie
                    order, truncated like datasets/dataloader.py:66-69
  * compare_pair()  all index lists of pair k inside the stacked batch == the pair's own lists (global row offsets and the
                    shadow index translated), points bit-equal
  * encoder_error() the stacked encoder output of pair k vs oracle/blocks_port.py on that pair alone (normwise)
Used by tests/test_gpu_benchsize.py and by bench.py's ``parity_check`` (outside every timed region).
"""
import numpy as np

import oracle


def _query(R_is_ref, q, s, ql, sl, radius, limit):
    if R_is_ref:
        rows, _ = oracle.ref_batch_query_canonical(q, s, ql, sl, radius, limit)
        return rows
    return oracle.port().batch_query(q, s, ql, sl, radius, limit)


def cpu_pyramid(src, tgt, limits, first_subsampling_dl, conv_radius, num_layers=4):
    use_ref = oracle.have_ref()
    R = oracle.ref() if use_ref else oracle.port()
    pts = np.concatenate([src, tgt]).astype(np.float32)
    lens = np.array([len(src), len(tgt)], np.int32)
    out = dict(points=[], neighbors=[], pools=[], upsamples=[], stack_lengths=[], kind="reference" if use_ref else "port")
    r = first_subsampling_dl * conv_radius
    for l in range(num_layers):
        out["neighbors"].append(_query(use_ref, pts, pts, lens, lens, r, limits[l]))
        if l + 1 < num_layers:
            pp, pl = R.subsample_batch(pts, lens, 2 * r / conv_radius)
            out["pools"].append(_query(use_ref, pp, pts, pl, lens, r, limits[l]))
            out["upsamples"].append(_query(use_ref, pts, pp, lens, pl, 2 * r, limits[l]))
        else:
            pp, pl = pts[:0], lens[:0]
            out["pools"].append(np.zeros((0, 1), np.int32))
            out["upsamples"].append(np.zeros((0, 1), np.int32))
        out["points"].append(pts)
        out["stack_lengths"].append(lens)
        pts, lens = pp, pl
        r *= 2
    return out


def _starts(lens):
    return np.concatenate([[0], np.cumsum(np.asarray(lens, np.int64))])


def compare_pair(batch, k, cpu):
    """batch: the stacked GPU pyramid (tensors); k: pair index; cpu: cpu_pyramid() of that pair.
    -> (number of arrays compared, list of mismatch descriptions)"""
    bad, n = [], 0
    L = len(cpu["points"])
    starts = [_starts(t.cpu().numpy()) for t in batch["stack_lengths"]]
    for l in range(L):
        s0, s1 = int(starts[l][2 * k]), int(starts[l][2 * k + 2])
        gp = batch["points"][l][s0:s1].cpu().numpy()
        n += 1
        if not (np.array_equal(batch["stack_lengths"][l][2 * k:2 * k + 2].cpu().numpy(), cpu["stack_lengths"][l])
                and np.array_equal(gp, cpu["points"][l])):
            bad.append(f"points level {l}")
            continue
        for name, ql, sl in (("neighbors", l, l), ("pools", l + 1, l), ("upsamples", l, l + 1)):
            if name != "neighbors" and l + 1 >= L:
                continue
            n += 1
            q0, q1 = int(starts[ql][2 * k]), int(starts[ql][2 * k + 2])
            so, ns_pair = int(starts[sl][2 * k]), int(starts[sl][2 * k + 2] - starts[sl][2 * k])
            ns_total = int(batch["points"][sl].shape[0])
            g = batch[name][l][q0:q1].cpu().numpy().astype(np.int64)
            c = cpu[name][l].astype(np.int64)
            c = np.where(c >= ns_pair, ns_total, c + so)
            w = g.shape[1]
            if c.shape[1] < w:          # the stacked width is min(limit, max_count of the WHOLE batch): pad with shadows
                c = np.concatenate([c, np.full((c.shape[0], w - c.shape[1]), ns_total, np.int64)], 1)
            if c.shape != g.shape or not np.array_equal(g, c):
                bad.append(f"{name} level {l}")
    return n, bad


def encoder_error(y_pair, cpu, state_dict, cfg, x=None):
    """y_pair: the GPU encoder's rows of one pair (coarsest level) -> normwise max error vs blocks_port on that pair.
    x: the pair's input feature rows (default: ones [N, in_feats_dim], the reference's geometry-only input)"""
    import torch
    from oracle import blocks_port as bp
    desc = bp.encoder_blocks_from_state_dict({k: v.cpu() for k, v in state_dict.items()}, prefix="encoder_blocks.",
                                             first_subsampling_dl=cfg.first_subsampling_dl, conv_radius=cfg.conv_radius,
                                             KP_extent=cfg.KP_extent)
    b = {k: [torch.from_numpy(np.ascontiguousarray(a)) for a in cpu[k]] for k in ("points", "neighbors", "pools", "upsamples")}
    with torch.no_grad():
        x0 = torch.ones(b["points"][0].shape[0], cfg.in_feats_dim) if x is None else torch.as_tensor(x)
        ref, _ = bp.encoder(x0, b, desc)
    y = y_pair.detach().cpu()
    return float((y - ref).abs().max() / ref.abs().max())
