"""faiss-shaped facade over the B200 engine.

Exactly the four names the reference uses from faiss (ref
minivectordb/vector_database.py:43-46, 475, 497, 511-514): ``IndexFlatIP``,
``.add``, ``.search``, ``normalize_L2`` (+ ``ntotal``).  With
``install_as_faiss()`` the reference's own modules and tests run on the GPU
engine unmodified (`import faiss` resolves to this module).
"""
from __future__ import annotations

import sys
import types

import numpy as np

from .engine import FlatIPEngine, normalize_L2  # noqa: F401  (re-exported)


class IndexFlatIP:
    def __init__(self, d: int, device: int = 0):
        self.d = int(d)
        self._engine = FlatIPEngine(self.d, device=device)

    @property
    def ntotal(self) -> int:
        return self._engine.ntotal

    def add(self, x) -> None:
        x = np.ascontiguousarray(x, dtype=np.float32)
        assert x.ndim == 2 and x.shape[1] == self.d  # faiss's SWIG wrapper asserts the same
        self._engine.add(x, normalize=False)

    def search(self, x, k: int):
        x = np.ascontiguousarray(x, dtype=np.float32)
        assert x.ndim == 2 and x.shape[1] == self.d
        assert k > 0
        return self._engine.search(x, int(k))

    def reset(self) -> None:
        self._engine.reset()


def install_as_faiss() -> types.ModuleType:
    m = types.ModuleType("faiss")
    m.IndexFlatIP = IndexFlatIP
    m.normalize_L2 = normalize_L2
    m.__mvdb_b200__ = True
    sys.modules["faiss"] = m
    return m
