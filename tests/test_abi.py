"""CPU: the C-ABI library loads and exports every symbol include/mvdb_b200.h declares."""
import ctypes
import os
import re

import pytest

from minivectordb_b200 import _native as N

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "mvdb_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mvdb_[a-zA-Z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert _declared_symbols() == sorted(N.EXPORTS)


def test_library_exports_every_declared_symbol():
    assert os.path.exists(N.LIB_PATH), "run __graft_entry__.build() first"
    L = ctypes.CDLL(N.LIB_PATH)
    for name in _declared_symbols():
        assert hasattr(L, name), f"{name} not exported"


def test_abi_version_and_error_channel():
    L = N.lib()
    assert L.mvdb_abi_version() == 1
    assert isinstance(L.mvdb_last_error(), bytes)


def test_no_cpu_fallback_without_device():
    if N.device_count() > 0:
        pytest.skip("a GPU is present")
    from minivectordb_b200 import FlatIPEngine
    with pytest.raises(N.MvdbError) as e:
        FlatIPEngine(16)
    assert e.value.code == N.MVDB_ERR_CUDA
    assert "no CPU fallback" in str(e.value)


def test_library_has_no_libcuda_link_dependency():
    # must dlopen on a GPU-less box: driver entry points are resolved at run time
    import subprocess
    out = subprocess.run(["ldd", N.LIB_PATH], capture_output=True, text=True).stdout
    assert "libcuda.so" not in out


def test_build_decision_follows_source_content_not_file_times(tmp_path, monkeypatch):
    """build() recompiles when (and only when) the sources or flags differ from what the shipped .so was
    built from -- a content hash next to the library, not mtimes."""
    assert os.path.exists(N.HASH_PATH), "run __graft_entry__.build() first"
    assert not N.needs_build()
    os.utime(os.path.join(N.SRC_DIR, "scan.cuh"))          # a newer mtime alone changes nothing
    assert not N.needs_build()
    monkeypatch.setattr(N, "NVCC_FLAGS", N.NVCC_FLAGS + ["-DX"])
    assert N.needs_build()                                   # different flags (or sources) do
