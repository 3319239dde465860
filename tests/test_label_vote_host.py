"""CPU: pins the features / classes branch of the grid subsampling (grid_subsampling.cpp:34-102).

* oracle/port.c (restatement) == the UNMODIFIED reference core (oracle/_ref) on feature means, class votes and their order;
* the DEVICE routine of the product (pcrcg_b200/csrc/label_vote.h, what k_label_vote runs per voxel) compiled here for the host
  by g++ == the live std::unordered_map<int,int> + max_element of the reference's libstdc++, including tied votes, hash
  collisions, negative labels and the rehash epochs (13 / 29 / 59 / 127 buckets).  Compiling that header for the host is test
  infrastructure: the product only ever runs it on the GPU.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")


@pytest.fixture(scope="module")
def host_vote(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("lv") / "liblabel_vote_host.so")
    subprocess.check_call(["g++", "-O2", "-fPIC", "-shared", "-std=c++17", "-o", so, os.path.join(ROOT, "tests", "host_shims", "label_vote_host.cpp")])
    L = C.CDLL(so)
    L.host_label_vote.restype = C.c_int
    L.host_label_vote.argtypes = [_i32p, C.c_long, C.POINTER(C.c_int)]

    def vote(labels):
        a = np.ascontiguousarray(labels, dtype=np.int32)
        ov = C.c_int(0)
        return int(L.host_label_vote(a, len(a), C.byref(ov))), bool(ov.value)
    return vote


def _label_sequences(rng, n_trials):
    for _ in range(n_trials):
        D = int(rng.integers(1, 65))
        kind = int(rng.integers(0, 5))
        if kind == 0:
            pool = rng.integers(-50, 50, size=D)
        elif kind == 1:
            pool = rng.integers(-2 ** 31, 2 ** 31 - 1, size=D)
        elif kind == 2:
            pool = np.arange(D) * 13 + rng.integers(0, 3)          # collisions in the 13-bucket table
        elif kind == 3:
            pool = rng.integers(0, 30, size=D)
        else:
            pool = rng.permutation(200)[:D] - 100                  # all distinct -> with n == D every vote ties
        n = D if kind == 4 else int(rng.integers(1, 4 * D + 2))
        lab = (rng.permutation(pool)[:n] if kind == 4 else rng.choice(pool, size=n)).astype(np.int32)
        if len(np.unique(lab)) <= 64:
            yield lab


def test_device_vote_routine_equals_live_unordered_map(host_vote, port):
    rng = np.random.default_rng(0)
    checker = oracle.ref() if oracle.have_ref() else port
    n = ties = deep = 0
    for lab in _label_sequences(rng, 6000):
        got, ov = host_vote(lab)
        assert not ov
        assert got == checker.label_vote(lab), lab.tolist()
        _, cnt = np.unique(lab, return_counts=True)
        n += 1
        ties += int((cnt == cnt.max()).sum() > 1)
        deep += int(len(cnt) > 13 and (cnt == cnt.max()).sum() > 1)
    assert n > 5000 and ties > 1500 and deep > 300, (n, ties, deep)    # the tie / rehash paths were really exercised


def test_device_vote_routine_reports_overflow(host_vote):
    lab = np.arange(65, dtype=np.int32)
    assert host_vote(lab)[1] and not host_vote(lab[:64])[1]


def test_port_vote_equals_reference(port):
    if not oracle.have_ref():
        pytest.skip("reference core not built here")
    rng = np.random.default_rng(1)
    for lab in _label_sequences(rng, 3000):
        assert port.label_vote(lab) == oracle.ref().label_vote(lab)


def _cloud_case(rng):
    nb = int(rng.integers(1, 4))
    lens = rng.integers(1, 400, size=nb).astype(np.int32)
    n = int(lens.sum())
    pts = (rng.random((n, 3)) * rng.choice([0.3, 1.0, 3.0])).astype(np.float32)
    fdim = int(rng.integers(1, 6))
    ldim = 1 if nb > 1 else int(rng.integers(1, 4))
    f = rng.standard_normal((n, fdim)).astype(np.float32)
    c = rng.integers(-3, rng.choice([2, 5, 40]), size=(n, ldim)).astype(np.int32)
    return pts, lens, f, c, float(rng.choice([0.05, 0.1, 0.3])), int(rng.choice([0, 0, 7]))


def test_port_features_classes_equal_reference(port):
    if not oracle.have_ref():
        pytest.skip("reference core not built here")
    rng = np.random.default_rng(2)
    for _ in range(120):
        pts, lens, f, c, dl, mp = _cloud_case(rng)
        for kw in (dict(features=f), dict(classes=c), dict(features=f, classes=c), dict(classes=c[:, 0])):
            a = port.subsample_batch_ex(pts, lens, sampleDl=dl, max_p=mp, **kw)
            b = oracle.ref().subsample_batch_ex(pts, lens, sampleDl=dl, max_p=mp, **kw)
            assert len(a) == len(b) == 2 + len(kw)
            for x, y in zip(a, b):
                assert x.dtype == y.dtype and np.array_equal(x, y)
        plain = port.subsample_batch(pts, lens, dl, mp)
        assert all(np.array_equal(x, y) for x, y in zip(plain, port.subsample_batch_ex(pts, lens, sampleDl=dl, max_p=mp)))


def test_multi_column_classes_need_a_single_cloud(port):
    pts = np.zeros((4, 3), np.float32)
    with pytest.raises(ValueError, match="157-158"):
        port.subsample_batch_ex(pts, [2, 2], classes=np.zeros((4, 2), np.int32))
