"""ctypes binding of the C ABI in include/mvdb_b200.h (libmvdb_b200.so).

This is the only door between the Python host layer and the CUDA engine; it
is what a maintainer of the reference would bind instead of `import faiss`
(ref minivectordb/vector_database.py:2).  There is no fallback: if the shared
library is missing or no CUDA device is visible, every compute call raises.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import threading

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
LIB_PATH = os.path.join(_PKG, "libmvdb_b200.so")
SRC_DIR = os.path.join(_PKG, "csrc")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared", "-diag-suppress", "1886",
]

MVDB_OK, MVDB_ERR_ARG, MVDB_ERR_CUDA, MVDB_ERR_OOM, MVDB_ERR_STATE = 0, -1, -2, -3, -4
SCAN_AUTO, SCAN_TMA, SCAN_LDG = 0, 1, 2

# every symbol include/mvdb_b200.h declares (checked by tests/test_abi.py)
EXPORTS = [
    "mvdb_abi_version", "mvdb_last_error", "mvdb_device_count", "mvdb_index_create",
    "mvdb_index_destroy", "mvdb_index_reset", "mvdb_index_set_option", "mvdb_index_add",
    "mvdb_index_add_device", "mvdb_index_add_synthetic", "mvdb_index_remove_rows",
    "mvdb_index_compact", "mvdb_index_dim", "mvdb_index_ntotal", "mvdb_index_reconstruct",
    "mvdb_index_reconstruct_n", "mvdb_index_device_view", "mvdb_index_workspace_create",
    "mvdb_index_workspace_destroy", "mvdb_index_search_device", "mvdb_index_search",
    "mvdb_normalize_L2", "mvdb_merge_topk_device", "mvdb_launch_count",
    "mvdb_exchange_create", "mvdb_exchange_ipc_handle", "mvdb_exchange_connect", "mvdb_exchange_set_offsets",
    "mvdb_exchange_status", "mvdb_exchange_destroy", "mvdb_index_search_exchange", "mvdb_debug_gemm_scores",
    "mvdb_index_mask_create", "mvdb_mask_destroy", "mvdb_index_search_with_mask",
    "mvdb_column_create", "mvdb_column_destroy", "mvdb_column_append", "mvdb_mask_from_predicate",
    "mvdb_mask_create_filled", "mvdb_mask_combine", "mvdb_mask_count", "mvdb_debug_read_trace", "mvdb_debug_read_gemm_prof",
    "mvdb_exchange_connect_local", "mvdb_exchange_set_option",
    "mvdb_group_create", "mvdb_group_destroy", "mvdb_group_set_option", "mvdb_group_search",
    "mvdb_debug_read_shadow_counters",
]


class MvdbError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"[mvdb rc={code}] {message}")
        self.code = code


def sources():
    return sorted(os.path.join(SRC_DIR, f) for f in os.listdir(SRC_DIR)
                  if f.endswith((".cu", ".cuh"))) + [os.path.join(_ROOT, "include", "mvdb_b200.h")]


HASH_PATH = LIB_PATH + ".srchash"


def source_hash() -> str:
    """Content hash of every source the library is built from plus the compiler flags: what decides
    whether the in-tree .so is current (file times do not survive a checkout or a snapshot)."""
    import hashlib
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for path in sources():
        h.update(os.path.basename(path).encode())
        with open(path, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH) or not os.path.exists(HASH_PATH):
        return True
    with open(HASH_PATH) as f:
        return f.read().strip() != source_hash()


def build(force: bool = False, verbose: bool = False) -> str:
    """nvcc -gencode arch=compute_100a,code=sm_100a ... -> libmvdb_b200.so (in-tree)."""
    if not force and not needs_build():
        return LIB_PATH
    nvcc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "bin", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + ["-ccbin", "/usr/bin/g++", "-o", LIB_PATH + ".tmp",
                                 os.path.join(SRC_DIR, "mvdb_b200.cu")]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    os.replace(LIB_PATH + ".tmp", LIB_PATH)
    with open(HASH_PATH, "w") as f:
        f.write(source_hash() + "\n")
    if verbose:
        print(res.stderr)
    return LIB_PATH


_lib = None
_lock = threading.Lock()

c_f32p = ctypes.POINTER(ctypes.c_float)
c_i64p = ctypes.POINTER(ctypes.c_int64)
c_u8p = ctypes.POINTER(ctypes.c_uint8)
c_u32p = ctypes.POINTER(ctypes.c_uint32)
c_vp = ctypes.c_void_p


def lib():
    """Load libmvdb_b200.so (never builds implicitly on a GPU box: the .so
    travels with the tree; build() is called by __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(the engine has no CPU fallback)")
        L = ctypes.CDLL(LIB_PATH)
        i, i64, u64 = ctypes.c_int, ctypes.c_int64, ctypes.c_uint64
        sig = {
            "mvdb_abi_version": (i, []),
            "mvdb_last_error": (ctypes.c_char_p, []),
            "mvdb_device_count": (i, [ctypes.POINTER(i)]),
            "mvdb_index_create": (i, [i, i, u64, ctypes.POINTER(c_vp)]),
            "mvdb_index_destroy": (i, [c_vp]),
            "mvdb_index_reset": (i, [c_vp]),
            "mvdb_index_set_option": (i, [c_vp, ctypes.c_char_p, i64]),
            "mvdb_index_add": (i, [c_vp, c_vp, u64, i, c_i64p]),
            "mvdb_index_add_device": (i, [c_vp, c_vp, u64, i, c_i64p]),
            "mvdb_index_add_synthetic": (i, [c_vp, u64, i64, u64, i, i, c_i64p]),
            "mvdb_index_remove_rows": (i, [c_vp, c_vp, u64]),
            "mvdb_index_compact": (i, [c_vp, c_i64p]),
            "mvdb_index_dim": (i, [c_vp, ctypes.POINTER(i)]),
            "mvdb_index_ntotal": (i, [c_vp, c_i64p, c_i64p]),
            "mvdb_index_reconstruct": (i, [c_vp, i64, c_vp]),
            "mvdb_index_reconstruct_n": (i, [c_vp, i64, u64, c_vp]),
            "mvdb_index_device_view": (i, [c_vp, ctypes.POINTER(c_vp), c_i64p, ctypes.POINTER(c_vp)]),
            "mvdb_index_workspace_create": (i, [c_vp, ctypes.POINTER(c_vp)]),
            "mvdb_index_workspace_destroy": (i, [c_vp]),
            "mvdb_index_search_device": (i, [c_vp, c_vp, c_vp, i64, i64, c_vp, u64, i, i64, c_vp, c_vp, c_vp]),
            "mvdb_index_search": (i, [c_vp, c_vp, i64, i64, c_vp, u64, i, c_vp, c_vp]),
            "mvdb_normalize_L2": (i, [c_vp, u64, i, i]),
            "mvdb_merge_topk_device": (i, [i, c_vp, c_vp, i, i64, i64, c_vp, c_vp, c_vp]),
            "mvdb_launch_count": (u64, []),
            "mvdb_exchange_create": (i, [i, i, i, i, i, ctypes.POINTER(c_vp)]),
            "mvdb_exchange_ipc_handle": (i, [c_vp, c_vp]),
            "mvdb_exchange_connect": (i, [c_vp, c_vp, c_vp]),
            "mvdb_exchange_set_offsets": (i, [c_vp, c_vp]),
            "mvdb_exchange_status": (i, [c_vp, ctypes.POINTER(i)]),
            "mvdb_exchange_destroy": (i, [c_vp]),
            "mvdb_debug_gemm_scores": (i, [c_vp, c_vp, i64, c_vp]),
            "mvdb_index_mask_create": (i, [c_vp, c_vp, u64, ctypes.POINTER(c_vp)]),
            "mvdb_mask_destroy": (i, [c_vp]),
            "mvdb_index_search_with_mask": (i, [c_vp, c_vp, i64, i64, c_vp, i, c_vp, c_vp]),
            "mvdb_column_create": (i, [c_vp, ctypes.POINTER(c_vp)]),
            "mvdb_column_destroy": (i, [c_vp]),
            "mvdb_column_append": (i, [c_vp, c_vp, c_vp, u64]),
            "mvdb_mask_from_predicate": (i, [c_vp, c_vp, i, ctypes.c_double, ctypes.POINTER(c_vp)]),
            "mvdb_mask_create_filled": (i, [c_vp, u64, ctypes.POINTER(c_vp)]),
            "mvdb_mask_combine": (i, [c_vp, c_vp, i]),
            "mvdb_mask_count": (i, [c_vp, ctypes.POINTER(u64)]),
            "mvdb_debug_read_trace": (i, [c_vp, c_vp]),
            "mvdb_debug_read_gemm_prof": (i, [c_vp, c_vp, ctypes.c_int]),
            "mvdb_index_search_exchange": (i, [c_vp, c_vp, c_vp, c_vp, i64, i64, c_vp, u64, i, c_vp, c_vp, c_vp]),
            "mvdb_exchange_connect_local": (i, [c_vp, i, c_vp]),
            "mvdb_exchange_set_option": (i, [c_vp, ctypes.c_char_p, i64]),
            "mvdb_group_create": (i, [c_vp, i, ctypes.POINTER(c_vp)]),
            "mvdb_group_destroy": (i, [c_vp]),
            "mvdb_group_set_option": (i, [c_vp, ctypes.c_char_p, i64]),
            "mvdb_group_search": (i, [c_vp, c_vp, i64, i64, c_vp, c_vp, c_vp, i, c_vp, c_vp]),
            "mvdb_debug_read_shadow_counters": (i, [c_vp, c_vp]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        if L.mvdb_abi_version() != 1:
            raise ImportError("libmvdb_b200.so ABI version mismatch; rebuild")
        _lib = L
        return _lib


def check(rc: int) -> None:
    if rc != MVDB_OK:
        msg = lib().mvdb_last_error()
        raise MvdbError(rc, msg.decode("utf-8", "replace") if msg else "unknown error")


def device_count() -> int:
    n = ctypes.c_int(0)
    check(lib().mvdb_device_count(ctypes.byref(n)))
    return n.value


def launch_count() -> int:
    return int(lib().mvdb_launch_count())
