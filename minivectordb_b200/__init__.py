"""minivectordb_b200 -- B200-native flat inner-product search behind MiniVectorDB's API.

Only the hot path of the reference (ref minivectordb/vector_database.py:466-536,
minivectordb/sharded_vector_database.py:598-662) is rebuilt here: an
HBM-resident matrix scanned by hand-written sm_100a kernels reached through the
C ABI in include/mvdb_b200.h.  See DESIGN.md.
"""
from .engine import DeviceColumn, FlatIPEngine, MaskHandle, ShardGroup, Workspace, normalize_L2, pack_mask, merge_topk_device  # noqa: F401
from . import faiss_shim  # noqa: F401

__all__ = ["DeviceColumn", "FlatIPEngine", "MaskHandle", "ShardGroup", "Workspace", "normalize_L2", "pack_mask", "merge_topk_device", "faiss_shim"]


def __getattr__(name):
    # the drop-in classes pull in sklearn; import them lazily
    if name == "VectorDatabase":
        from .vector_database import VectorDatabase
        return VectorDatabase
    if name == "ShardedVectorDatabase":
        from .sharded_vector_database import ShardedVectorDatabase
        return ShardedVectorDatabase
    raise AttributeError(name)
