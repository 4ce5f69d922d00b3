"""The C-ABI library loads and exports every symbol include/mvs_b200.h declares (no GPU needed)."""
import ctypes
import os
import re

import pytest

import mvs_b200
from mvs_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "mvs_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mvsb200_\w+)\s*\(", src)))


def test_header_symbols_are_exported():
    names = _declared()
    assert len(names) >= 12
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/mvs_b200.h but not exported"


def test_binding_covers_header():
    assert sorted(_lib.exported_symbols()) == _declared()


def test_abi_version_and_error_string():
    lib = mvs_b200.load_library()
    assert lib.mvsb200_abi_version() == _lib.ABI_VERSION
    # argument validation happens before any CUDA call, so it can be exercised without a GPU
    rc = lib.mvsb200_warp_variance_fwd(None, None, None, None, 0, 1, 3, 32, 8, 4, 4, None)
    assert rc == -1 and b"null pointer" in lib.mvsb200_last_error()
    rc = lib.mvsb200_softmax_depth_fwd(None, 1, None, None, None, None, 1, 8, 4, 4, 5, None)
    assert rc == -1


def test_ops_refuse_cpu_tensors():
    import torch
    with pytest.raises(mvs_b200.MvsB200Error):
        mvs_b200.extract_depth_map(torch.rand(1, 1, 8, 4, 4), torch.rand(1, 8, 1, 1))
    with pytest.raises(mvs_b200.MvsB200Error):
        mvs_b200.softmax_over_depth(torch.rand(1, 1, 8, 4, 4))


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libmvs_b200.so")
    with pytest.raises(mvs_b200.MvsB200Error, match="no CPU fallback"):
        _lib.load()
