"""CPU: the thread bodies of the three kernels behind subsample_batch(features=, classes=) -- pcrcg_b200/csrc/subsample_extras.h,
the code nvcc compiles into k_bary_feat / k_label_vote / k_gather_extra -- compiled for the host and run over the same launch
geometry, against the CPU oracle.  The workspace arrays the kernels read (sorted voxel runs, first-occurrence ranks, the final
lists k_order leaves in seqA / seqB) are rebuilt here with NumPy from the oracle's voxel keys and output order, with voxel slots in
a random order as the GPU's hash table produces them.  What this does NOT cover: the plain pipeline that produces those arrays on
the GPU (bit-exact in tests/test_gpu_preprocess.py) and the launch plumbing (tests/test_zz_gpu_round2_late.py)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCHED = np.array([13, 29, 59, 127, 257, 541, 1109, 2357, 5087, 10273, 20753, 42043, 85229, 172933, 351061, 712697], np.uint32)


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("sx") / "libsubsample_extras_host.so")
    subprocess.check_call(["g++", "-O2", "-fPIC", "-shared", "-std=c++17", "-ffp-contract=off", "-o", so,
                           os.path.join(ROOT, "tests", "host_shims", "subsample_extras_host.cpp")])
    L = C.CDLL(so)
    L.host_subsample_extras.restype = C.c_int
    return L


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _workspace(port, pts, lens, dl, max_p, rng):
    """What subsample_batch_dev leaves behind for a stacked batch, rebuilt on the CPU."""
    n, nb = len(pts), len(lens)
    starts = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    slot = np.empty(n, np.int64)
    base = 0
    for c in range(nb):
        p = pts[starts[c]:starts[c + 1]]
        keys, _, _ = port.voxel_keys(p, dl)
        uniq, inv = np.unique(keys, return_inverse=True)
        perm = rng.permutation(len(uniq))                       # hash-table slots: an arbitrary order of the voxels
        slot[starts[c]:starts[c + 1]] = base + perm[inv]
        base += 2 * len(p) + 1                                  # the cloud's table region
    sidx = np.argsort(slot, kind="stable").astype(np.uint32)
    sslot = slot[sidx].astype(np.uint32)
    head = np.ones(n, bool)
    head[1:] = sslot[1:] != sslot[:-1]
    flag = np.zeros(n + 1, np.uint32)
    flag[sidx[head]] = 1                                        # k_heads: the run head is the voxel's first occurrence
    rank = np.concatenate([[0], np.cumsum(flag[:n])]).astype(np.uint32)
    # voxel rank (within its cloud) of every point: rank of the voxel's first point minus the cloud's base
    first_of_run = np.maximum.accumulate(np.where(head, np.arange(n), 0))
    vrank = np.empty(n, np.int64)
    vrank[sidx] = rank[sidx[first_of_run]]
    cloud = np.repeat(np.arange(nb), lens)
    label = (vrank - rank[starts[cloud]]).astype(np.int32)
    # the final list of every cloud = the voxel rank at every output position: a class column that is constant per voxel comes
    # back from the oracle in output order
    _, full_lens, order = port.subsample_batch_ex(pts, lens, classes=label, sampleDl=dl)
    out_lens = np.minimum(full_lens, max_p).astype(np.int32) if max_p > 0 else full_lens.astype(np.int32)
    out_base = np.concatenate([[0], np.cumsum(out_lens)]).astype(np.int32)
    U = int(rank[n])
    seqA = np.full(max(U, 1), 0xdeadbeef, np.uint32)
    seqB = np.full(max(U, 1), 0xdeadbeef, np.uint32)
    o = 0
    for c in range(nb):
        M, Ub = int(full_lens[c]), int(rank[starts[c]])
        assert M == int(rank[starts[c + 1]]) - Ub
        epochs, done = 0, 0
        while done < M:
            done = min(M, int(SCHED[epochs]))
            epochs += 1
        (seqB if epochs & 1 else seqA)[Ub:Ub + M] = order[o:o + M, 0]
        o += M
    return dict(starts=starts, sslot=sslot, sidx=sidx, rank=rank, out_lens=out_lens, out_base=out_base, seqA=seqA, seqB=seqB, U=U)


def _run(shim, ws, n, nb, f, c):
    fdim = 0 if f is None else f.shape[1]
    ldim = 0 if c is None else c.shape[1]
    featU = np.full((max(ws["U"], 1), max(fdim, 1)), np.nan, np.float32)
    clsU = np.full((max(ws["U"], 1), max(ldim, 1)), -777, np.int32)
    m = int(ws["out_lens"].sum())
    of = np.full((max(m, 1), max(fdim, 1)), np.nan, np.float32)
    oc = np.full((max(m, 1), max(ldim, 1)), -777, np.int32)
    ov = shim.host_subsample_extras(n, nb, _ptr(f), fdim, _ptr(c), ldim, _ptr(ws["sslot"]), _ptr(ws["sidx"]), _ptr(ws["rank"]),
                                    _ptr(ws["starts"]), _ptr(ws["out_lens"]), _ptr(ws["out_base"]), _ptr(ws["seqA"]), _ptr(ws["seqB"]),
                                    _ptr(SCHED), _ptr(featU), _ptr(clsU), _ptr(of), _ptr(oc))
    return ov, of[:m, :fdim], oc[:m, :ldim]


@pytest.mark.parametrize("seed", range(12))
def test_kernel_bodies_match_the_oracle(shim, port, seed):
    rng = np.random.default_rng(seed)
    nb = int(rng.integers(1, 5))
    lens = rng.integers(1, 900, size=nb).astype(np.int32)
    n = int(lens.sum())
    extent, dl = [(1.0, 0.1), (0.3, 0.1), (2.0, 0.05), (1.0, 0.31)][seed % 4]        # 0.3 / 0.1: crowded voxels (deep label sets)
    pts = (rng.random((n, 3)) * extent).astype(np.float32)
    fdim = int(rng.choice([1, 3, 70]))                                                # 70 > 64: the column stride of blockIdx.y
    ldim = 1 if nb > 1 else int(rng.choice([1, 2, 66]))
    f = rng.standard_normal((n, fdim)).astype(np.float32)
    c = rng.integers(-3, rng.choice([2, 6, 40]), size=(n, ldim)).astype(np.int32)
    max_p = int(rng.choice([0, 0, 23]))
    ws = _workspace(port, pts, lens, dl, max_p, rng)
    want = port.subsample_batch_ex(pts, lens, features=f, classes=c, sampleDl=dl, max_p=max_p)
    assert np.array_equal(want[1], ws["out_lens"])
    ov, of, oc = _run(shim, ws, n, nb, f, c)
    assert ov == 0
    assert np.array_equal(of, want[2]) and np.array_equal(oc, want[3])
    ov, of, _ = _run(shim, ws, n, nb, f, None)                                        # features only / classes only
    assert np.array_equal(of, want[2])
    ov, _, oc = _run(shim, ws, n, nb, None, c)
    assert np.array_equal(oc, want[3])


def test_rehash_epochs_select_the_right_list(shim, port):
    """clouds whose voxel counts sit on both sides of 13 / 29 / 59 / 127: the final list alternates between seqA and seqB"""
    rng = np.random.default_rng(99)
    for M in (1, 12, 13, 14, 29, 30, 59, 60, 127, 128, 300):
        g = int(np.ceil(M ** (1 / 3))) + 1
        cells = rng.permutation(g ** 3)[:M]
        centres = np.stack([cells % g, (cells // g) % g, cells // (g * g)], 1).astype(np.float32) + 0.5
        pts = np.repeat(centres, 3, axis=0) + rng.uniform(-0.3, 0.3, size=(3 * M, 3)).astype(np.float32)
        pts = np.concatenate([pts, np.zeros((1, 3), np.float32) + 0.5]) if 0 in cells else np.concatenate([pts, centres[:1]])
        pts = pts[rng.permutation(len(pts))].astype(np.float32)
        lens = np.array([len(pts)], np.int32)
        f = rng.standard_normal((len(pts), 2)).astype(np.float32)
        c = rng.integers(0, 3, size=(len(pts), 1)).astype(np.int32)
        ws = _workspace(port, pts, lens, 1.0, 0, rng)
        want = port.subsample_batch_ex(pts, lens, features=f, classes=c, sampleDl=1.0)
        _, of, oc = _run(shim, ws, len(pts), 1, f, c)
        assert np.array_equal(of, want[2]) and np.array_equal(oc, want[3]), M


def test_overflow_is_reported(shim, port):
    rng = np.random.default_rng(5)
    pts = (rng.random((70, 3)) * 0.01).astype(np.float32)
    lens = np.array([70], np.int32)
    ws = _workspace(port, pts, lens, 1.0, 0, rng)
    ov, _, _ = _run(shim, ws, 70, 1, None, np.arange(70, dtype=np.int32).reshape(-1, 1))
    assert ov == 1
    ov, _, _ = _run(shim, ws, 70, 1, None, (np.arange(70, dtype=np.int32) % 64).reshape(-1, 1))
    assert ov == 0
