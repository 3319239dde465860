"""CPU: the C-ABI library loads, exports every symbol include/pcrcg_b200.h declares, and the Python
binding table matches the header.  No compute calls (no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "pcrcg_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pcrcg_[a-z0-9_]+)\s*\(", src)) - {"pcrcg_stream_t"})


def test_header_symbols_exported():
    import __graft_entry__ as g
    if not os.path.exists(g.LIB):
        g.build()
    lib = ctypes.CDLL(g.LIB)
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/pcrcg_b200.h but not exported"


def test_binding_table_matches_header():
    from pcrcg_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared()
    L = _lib.lib()
    assert L.pcrcg_version() >= 100
    assert L.pcrcg_subsample_ws_bytes(1000, 2) > 0 and L.pcrcg_radius_ws_bytes(1000, 1000, 2) > 0
    assert L.pcrcg_profile_classes() >= 10 and L.pcrcg_profile_class_name(8) == b"kpconv_fused" and L.pcrcg_profile_class_name(9) == b"linear" and L.pcrcg_profile_class_name(4) == b"gemm"


def test_no_oracle_in_product_path():
    """The product package must never import the CPU checker."""
    pkg = os.path.join(ROOT, "pcrcg_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f"{f} imports oracle"


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from pcrcg_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.lib()


def test_ops_refuse_cpu_tensors():
    import torch
    from pcrcg_b200 import ops
    with pytest.raises(RuntimeError, match="CUDA tensors required"):
        ops.subsample_batch(torch.zeros(4, 3), torch.tensor([4], dtype=torch.int32), 0.1)
    with pytest.raises(RuntimeError, match="CUDA tensors required"):
        ops.kpconv_forward(torch.zeros(4, 3), torch.zeros(4, 3), torch.zeros(4, 2, dtype=torch.int64), torch.zeros(4, 1),
                           torch.zeros(15, 3), torch.zeros(15, 1, 8), 0.05)


def test_features_and_classes_argument_checks():
    """subsample_batch(features=, classes=): the reference's argument errors (wrapper.cpp:176-246) are raised before any device
    call, and the one case the reference leaves undefined is refused (no GPU needed)"""
    import numpy as np
    from pcrcg_b200.cpp_wrappers.cpp_subsampling import grid_subsampling as cpp_subsampling
    pts = np.zeros((4, 3), np.float32)
    with pytest.raises(RuntimeError, match=r"features.shape is not \(N, d\)"):
        cpp_subsampling.subsample_batch(pts, [4], features=np.ones(4, np.float32))
    with pytest.raises(RuntimeError, match=r"features.shape is not \(N, d\)"):
        cpp_subsampling.subsample_batch(pts, [4], features=np.ones((3, 2), np.float32))
    with pytest.raises(RuntimeError, match=r"classes.shape is not \(N,\) or \(N, d\)"):
        cpp_subsampling.subsample_batch(pts, [4], classes=np.zeros((4, 1, 1), np.int32))
    with pytest.raises(RuntimeError, match=r"classes.shape is not \(N,\) or \(N, d\)"):
        cpp_subsampling.subsample_batch(pts, [4], classes=np.zeros(5, np.int32))
    with pytest.raises(RuntimeError, match="single cloud only"):
        cpp_subsampling.subsample_batch(pts, [2, 2], classes=np.zeros((4, 2), np.int32))
    with pytest.raises(RuntimeError, match="Error converting input features"):
        cpp_subsampling.subsample_batch(pts, [4], features=[["a", "b"]] * 4)
    with pytest.raises(RuntimeError, match="Error parsing method"):
        cpp_subsampling.subsample_batch(pts, [4], method="nonsense")


def test_header_is_plain_c_and_a_c_caller_fails_loudly_without_a_gpu(tmp_path):
    """include/pcrcg_b200.h compiles as C99 and as C++11; examples/c_abi_demo.c links against the library and, on a box without a
    CUDA device, gets a non-zero status and a message from pcrcg_last_error() (no crash, no CPU fallback).  On a GPU box it runs."""
    import subprocess
    import __graft_entry__ as g
    if not os.path.exists(g.LIB):
        g.build()
    inc = os.path.join(ROOT, "include")
    cpp = tmp_path / "h.cpp"
    cpp.write_text('#include "pcrcg_b200.h"\n')
    subprocess.check_call(["g++", "-std=c++11", "-Wall", "-pedantic", "-Werror", "-I", inc, "-c", str(cpp), "-o", str(tmp_path / "h.o")])
    exe = str(tmp_path / "c_abi_demo")
    libdir = os.path.dirname(g.LIB)
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", inc, os.path.join(ROOT, "examples", "c_abi_demo.c"),
                           "-L", libdir, "-lpcrcg_b200", f"-Wl,-rpath,{libdir}", "-o", exe])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert "libpcrcg_b200 version 100" in r.stdout
    import torch
    if torch.cuda.is_available():
        assert r.returncode == 0 and "voxels" in r.stdout
    else:
        assert r.returncode == 2 and "pcrcg_subsample_batch_host failed:" in r.stdout and "cudaMalloc" in r.stdout


def test_host_entry_points_marshal_and_fail_loudly_without_a_gpu():
    """On a box without a CUDA device the host-buffer entry points are still CALLED (every argument marshalled through the binding
    table: a wrong count or type would raise ctypes.ArgumentError) and answer with the library's error, not with a result."""
    import numpy as np
    import torch
    if torch.cuda.is_available():
        pytest.skip("covered by the GPU tests")
    from pcrcg_b200.cpp_wrappers.cpp_subsampling import grid_subsampling as cpp_subsampling
    from pcrcg_b200.cpp_wrappers.cpp_neighbors import radius_neighbors as cpp_neighbors
    pts = np.random.default_rng(0).random((50, 3)).astype(np.float32)
    for kw in ({}, dict(features=np.ones((50, 2), np.float32)), dict(classes=np.zeros(50, np.int32)),
               dict(features=np.ones((50, 2), np.float32), classes=np.zeros((50, 3), np.int32))):
        with pytest.raises(RuntimeError, match="cudaMalloc|CUDA|cuda"):
            cpp_subsampling.subsample_batch(pts, [50], sampleDl=0.1, **kw)
    with pytest.raises(RuntimeError, match="cudaMalloc|CUDA|cuda"):
        cpp_neighbors.batch_query(pts, pts, [50], [50], radius=0.2)


def test_subsample_ex_device_entry_marshals_and_checks_its_arguments():
    from pcrcg_b200 import _lib
    L = _lib.lib()
    base, ex = L.pcrcg_subsample_ws_bytes(1000, 2), L.pcrcg_subsample_ex_ws_bytes(1000, 2, 4, 1)
    assert ex >= base + 1000 * 4 * 4 + 1000 * 4 and L.pcrcg_subsample_ex_ws_bytes(1000, 2, 0, 0) >= base
    rc = L.pcrcg_subsample_batch_ex_dev(None, 10, None, 1, 0.1, 0, None, 0, None, 0, None, None, None, None, None, None, 0, None)
    assert rc != 0 and b"workspace too small" in L.pcrcg_last_error()
    rc = L.pcrcg_subsample_batch_ex_dev(None, 10, None, 2, 0.1, 0, None, 0, 1, 2, None, None, None, 1, 1, None, 0, None)
    assert rc != 0 and b"single cloud only" in L.pcrcg_last_error()      # classes with 2 columns and 2 clouds
    rc = L.pcrcg_subsample_batch_ex_dev(None, 10, None, 1, 0.1, 0, 1, 3, None, 0, None, None, None, None, None, None, 0, None)
    assert rc != 0 and b"features need" in L.pcrcg_last_error()          # features without an output buffer
