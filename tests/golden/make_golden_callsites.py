"""Golden for the reference's own CALL SITES of the hot path: datasets/dataloader.py:14-69 (batch_grid_subsampling_kpconv,
batch_neighbors_kpconv) and :203-400 (collate_fn_descriptor) are taken from the reference's SOURCE with ast and executed unchanged
(the module itself cannot be imported: it loads the cp37 cpp_wrappers binaries and dataset modules that need open3d).  The two
extension modules they call, `cpp_subsampling` / `cpp_neighbors`, are bound to the UNMODIFIED reference C++ core (oracle/_ref)
behind a restatement of the CPython glue's tuple conventions (wrapper.cpp:318-326, cpp_neighbors/wrapper.cpp:211-227); neighbour
rows are put into the canonical (d2, index) tie order before the reference's own code truncates them.

Needs /root/reference; never runs on the GPU box.  Re-run:  python tests/golden/make_golden_callsites.py
-> tests/golden/callsites_ref.npz (inputs + the reference's outputs), checked on the GPU by tests/test_zz_gpu_round2_late.py.
"""
import ast
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True

import oracle                                     # noqa: E402
from pcrcg_b200 import synthetic                  # noqa: E402

N_CANON = [0]


def _np(a, dtype):
    if torch.is_tensor(a):
        a = a.detach().cpu().numpy()
    return np.ascontiguousarray(np.asarray(a), dtype=dtype)


class cpp_subsampling:                            # noqa: N801  (the reference's module alias)
    @staticmethod
    def subsample_batch(points, batches, features=None, classes=None, sampleDl=0.1, method="barycenters", max_p=0, verbose=0):
        f = None if features is None else _np(features, np.float32)
        c = None if classes is None else _np(classes, np.int32)
        return oracle.ref().subsample_batch_ex(_np(points, np.float32), _np(batches, np.int32), features=f, classes=c, sampleDl=sampleDl,
                                               max_p=max_p)


class cpp_neighbors:                              # noqa: N801
    @staticmethod
    def batch_query(queries, supports, q_batches, s_batches, radius=0.1):
        q, s = _np(queries, np.float32), _np(supports, np.float32)
        raw = oracle.ref().batch_query(q, s, _np(q_batches, np.int32), _np(s_batches, np.int32), radius)
        rows, changed = oracle.port().canonicalise_rows(q, s, raw)
        N_CANON[0] += changed
        return rows


class Cfg(dict):
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


def main():
    tree = ast.parse(open(f"{REF}/datasets/dataloader.py").read())
    wanted = ("batch_grid_subsampling_kpconv", "batch_neighbors_kpconv", "square_distance", "point2node", "point2node_correspondences",
              "collate_fn_descriptor")
    ns = {"torch": torch, "np": np, "cpp_subsampling": cpp_subsampling, "cpp_neighbors": cpp_neighbors}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in wanted:
            exec(compile(ast.Module(body=[node], type_ignores=[]), "datasets/dataloader.py", "exec"), ns)
    g = {}
    rng = np.random.default_rng(17)

    # ---- :14-52 batch_grid_subsampling_kpconv, all four branches, CPU tensors in as collate passes them ----------------------
    src, tgt, _ = synthetic.match3d_pair(17, n_target=2500)
    pts = np.concatenate([src, tgt]).astype(np.float32)
    lens = np.array([len(src), len(tgt)], np.int32)
    feats = rng.standard_normal((len(pts), 4)).astype(np.float32)
    labels = rng.integers(0, 5, size=(len(pts), 1)).astype(np.int32)
    P, B = torch.from_numpy(pts), torch.from_numpy(lens)
    g.update(sub_points=pts, sub_lens=lens, sub_features=feats, sub_labels=labels)
    fn = ns["batch_grid_subsampling_kpconv"]
    for tag, kw in (("plain", {}), ("feat", dict(features=torch.from_numpy(feats))), ("lab", dict(labels=torch.from_numpy(labels))),
                    ("both", dict(features=torch.from_numpy(feats), labels=torch.from_numpy(labels)))):
        out = fn(P, B, sampleDl=0.06, **kw)
        for i, o in enumerate(out):
            g[f"sub_{tag}_{i}"] = o.numpy()
    out = fn(P, B, sampleDl=0.06, max_p=300)
    g["sub_maxp_0"], g["sub_maxp_1"] = out[0].numpy(), out[1].numpy()

    # ---- :54-69 batch_neighbors_kpconv: truncated and full width ---------------------------------------------------------------
    sp, sl = g["sub_plain_0"], g["sub_plain_1"]
    fnn = ns["batch_neighbors_kpconv"]
    g["nb_pool_20"] = fnn(torch.from_numpy(sp), P, torch.from_numpy(sl), B, 0.15, 20).numpy()
    g["nb_conv_full"] = fnn(torch.from_numpy(sp), torch.from_numpy(sp), torch.from_numpy(sl), torch.from_numpy(sl), 0.15, 0).numpy()

    # ---- :203-400 collate_fn_descriptor on one pair ------------------------------------------------------------------------------
    sys.path.insert(0, REF)
    from configs.models import architectures
    cfg = Cfg(num_layers=4, first_subsampling_dl=0.025, conv_radius=2.5, deform_radius=5.0, architecture=architectures["indoor"])
    limits = [34, 39, 39, 38]
    corr = torch.from_numpy(np.stack([rng.integers(0, len(src), 600), rng.integers(0, len(tgt), 600)], 1))
    item = dict(src_pcd=src.astype(np.float32), tgt_pcd=tgt.astype(np.float32), src_feats=np.ones((len(src), 1), np.float32),
                tgt_feats=np.ones((len(tgt), 1), np.float32), rot=np.eye(3, dtype=np.float32), trans=np.zeros((3, 1), np.float32),
                correspondences=corr, sample="synthetic@17")
    with torch.no_grad():
        d = ns["collate_fn_descriptor"]([item], cfg, limits)
    g.update(col_src=item["src_pcd"], col_tgt=item["tgt_pcd"], col_corr=corr.numpy(), col_limits=np.array(limits, np.int32))
    for k in ("points", "neighbors", "pools", "upsamples", "stack_lengths"):
        for l, t in enumerate(d[k]):
            g[f"col_{k}_{l}"] = t.numpy()
    g["col_features"] = d["features"].numpy()
    g["col_node_overlap_gt"] = d["node_overlap_gt"].numpy()
    g["col_points2node"] = d["points2node"].numpy()
    g["col_keys"] = np.array(sorted(d.keys()))
    g["rows_canonicalised"] = np.int64(N_CANON[0])
    out = os.path.join(ROOT, "tests", "golden", "callsites_ref.npz")
    np.savez_compressed(out, **g)
    print("wrote", out, os.path.getsize(out), "bytes; rows canonicalised:", N_CANON[0])
    print("collate keys:", list(g["col_keys"]))
    print("levels:", [g[f"col_points_{l}"].shape[0] for l in range(4)], "widths:", [g[f"col_neighbors_{l}"].shape[1] for l in range(4)])


if __name__ == "__main__":
    main()
