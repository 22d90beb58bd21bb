"""CPU-only: the C-ABI libraries load and export every symbol include/*.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "chatllm.cpp_b200", "lib")


def _declared(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b((?:b200|ggml_backend)_[a-z0-9_]+)\s*\(", src)))


def test_kernel_abi_exports():
    path = os.path.join(LIBDIR, "libchatllm_b200.so")
    if not os.path.exists(path):
        pytest.skip("libchatllm_b200.so not built (run __graft_entry__.build())")
    L = ctypes.CDLL(path)
    names = _declared("chatllm_b200.h")
    assert len(names) >= 14
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert L.b200_abi_version() == 1


def test_python_mirror_lists_every_export():
    import __graft_entry__ as ge
    pkg = ge.load_package()
    assert sorted(pkg.EXPORTS) == _declared("chatllm_b200.h")


def test_plugin_exports():
    path = os.path.join(LIBDIR, "libggml-cuda.so")
    hdr = os.path.join(ROOT, "include", "ggml_b200_backend.h")
    if not (os.path.exists(path) and os.path.exists(hdr)):
        pytest.skip("plugin not built")
    base = os.path.join(ROOT, "oracle", "_ref", "lib", "libggml-base.so")
    if not os.path.exists(base):
        pytest.skip("host SDK library (oracle/_ref/lib/libggml-base.so) not present")
    ctypes.CDLL(base, mode=ctypes.RTLD_GLOBAL)
    L = ctypes.CDLL(path)
    for n in _declared("ggml_b200_backend.h"):
        assert hasattr(L, n), n
