"""The C-ABI shared library builds for sm_100a, loads, and exports every symbol include/rlap_b200.h
declares (no compute calls: this runs without a GPU)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "rlap_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rlap_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_expected_surface():
    names = _declared()
    for must in ("rlap_ingest", "rlap_schur_eliminate", "rlap_schur_emit", "rlap_approximate_cholesky_host",
                 "rlap_ingest_workspace_bytes", "rlap_schur_workspace_bytes"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from rlap_b200 import _build, _native
    if not os.path.exists(_build.LIB_PATH):
        _build.build()
    lib = ctypes.CDLL(_build.LIB_PATH)
    for name in _declared():
        assert hasattr(lib, name), name
    assert sorted(_native.EXPORTS) == _declared()
    assert _native.lib().rlap_version() >= 1
    assert _native.lib().rlap_status_string(4) == b"adjacency matrix is not symmetric"


def test_library_contains_sm_100a_code():
    from rlap_b200 import _build
    out = subprocess.run(["cuobjdump", "--list-elf", _build.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in out.stdout


def test_workspace_queries_validate_arguments():
    from rlap_b200 import _native
    L = _native.lib()
    b = ctypes.c_size_t(0)
    assert L.rlap_ingest_workspace_bytes(1000, 5000, ctypes.byref(b)) == 0 and b.value > 5000 * 8
    assert L.rlap_ingest_workspace_bytes(-1, 5, ctypes.byref(b)) == 1
    assert L.rlap_schur_workspace_bytes(1000, 5000, 1, 4, 0, 0, 0, ctypes.byref(b)) == 0
    one = b.value
    assert L.rlap_schur_workspace_bytes(1000, 5000, 1, 8, 0, 0, 0, ctypes.byref(b)) == 0 and b.value > one
    assert L.rlap_schur_workspace_bytes(1000, 5000, 0, 4, 0, 0, 0, ctypes.byref(b)) == 1
    assert L.rlap_schur_workspace_bytes(1 << 29, 5000, 1, 4, 0, 0, 0, ctypes.byref(b)) == 1   # V*n too large


def test_no_cpu_fallback():
    """the product path must fail loudly without a GPU"""
    import torch
    import rlap_b200
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    ei = torch.tensor([[0, 1], [1, 0]])
    with pytest.raises(RuntimeError):
        rlap_b200.ops.approximate_cholesky(ei, None, 2, 1, "random", "asc")


def test_product_does_not_touch_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "rlap_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "from oracle" not in txt and "import oracle" not in txt and "oracle/" not in txt.replace(
                    "oracle/rlap_oracle.cc (keyed mode)", ""), os.path.join(dirpath, f)


def test_expand_cols_host():
    """the host half of the column-pointer output needs no GPU: col rebuilt from column pointers, any thread count,
    empty views and empty columns included"""
    import numpy as np
    import torch
    from rlap_b200 import ops
    rng = np.random.default_rng(0)
    V, n = 5, 1000
    cnt = rng.integers(0, 6, size=(V, n))
    cnt[:, :50] += rng.integers(0, 400, size=(V, 50))
    cnt[2] = 0
    cp = np.zeros((V, n + 1), dtype=np.int32)
    cp[:, 1:] = np.cumsum(cnt, axis=1)
    vp = np.concatenate([[0], np.cumsum(cp[:, -1])]).astype(np.int64)
    ref = np.concatenate([np.repeat(np.arange(n, dtype=np.int32), cnt[v]) for v in range(V)])
    for threads in (1, 3, 8, 64):
        out = ops.expand_cols(torch.from_numpy(cp), torch.from_numpy(vp), threads=threads).numpy()
        assert np.array_equal(out, ref), threads
