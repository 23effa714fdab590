"""FlatIPEngine -- Python handle on one HBM-resident flat inner-product index.

Host-side mirror of the interface the reference binds from faiss
(ref minivectordb/vector_database.py:43-46, 475, 497, 511-514):

    faiss.IndexFlatIP(d)   -> FlatIPEngine(d)            / faiss_shim.IndexFlatIP(d)
    index.add(x)           -> engine.add(x)
    index.search(q, k)     -> engine.search(q, k)        (+ mask=, the filtered branch)
    faiss.normalize_L2(x)  -> normalize_L2(x)

plus what the numpy matrix gave the reference for free: row fetch
(`self.embeddings[row]`, VDB:55 -> reconstruct), row deletion (np.delete,
VDB:126 -> remove_rows + compact).  All arithmetic happens in libmvdb_b200.so
on the GPU; there is no CPU path.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Tuple

import numpy as np

from . import _native as N


def _as_f32_2d(a, d: Optional[int] = None, what: str = "array") -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.float32)
    if a.ndim == 1:
        a = a[None, :]
    if a.ndim != 2:
        raise ValueError(f"{what} must be 2-D (rows x d)")
    if d is not None and a.shape[1] != d:
        # faiss's SWIG wrapper asserts `d == self.d`; surface it as ValueError
        raise ValueError(f"{what} has dimension {a.shape[1]}, index has {d}")
    return a


def pack_mask(admissible: np.ndarray) -> np.ndarray:
    """bool[n] -> bytes for the engine (bit r&7 of byte r>>3 = row r admissible)."""
    return np.packbits(np.asarray(admissible, dtype=bool), bitorder="little")


class MaskHandle:
    """A filter bitmask kept in HBM (mvdb_index_mask_create / mvdb_mask_from_predicate).  Pass it
    as `mask=` to FlatIPEngine.search: no mask bytes move per query, and concurrent single-query
    searches carrying handles are coalesced into one tensor-core batch with per-query filters."""

    def __init__(self, engine: "FlatIPEngine", admissible=None, _raw=None, _rows=0):
        self._engine = engine
        self._h = ctypes.c_void_p()
        if _raw is not None:
            self._h, self.rows = _raw, int(_rows)
            return
        a = np.asarray(admissible)
        if a.dtype != np.bool_:
            raise TypeError("MaskHandle needs a bool[n] array of admissible rows")
        self.rows = int(a.shape[0])
        packed = pack_mask(a)
        N.check(N.lib().mvdb_index_mask_create(engine.handle, packed.ctypes.data if packed.size else None, self.rows,
                                               ctypes.byref(self._h)))

    # device-side combinators (in place): and / or / and-not
    def iand(self, other: "MaskHandle") -> "MaskHandle":
        N.check(N.lib().mvdb_mask_combine(self._h, other._h, 0))
        return self

    def ior(self, other: "MaskHandle") -> "MaskHandle":
        N.check(N.lib().mvdb_mask_combine(self._h, other._h, 1))
        return self

    def iandnot(self, other: "MaskHandle") -> "MaskHandle":
        N.check(N.lib().mvdb_mask_combine(self._h, other._h, 2))
        return self

    def count(self) -> int:
        """Admissible rows that are still live."""
        c = ctypes.c_uint64(0)
        N.check(N.lib().mvdb_mask_count(self._h, ctypes.byref(c)))
        return int(c.value)

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            N.lib().mvdb_mask_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DeviceColumn:
    """A numeric metadata column resident in HBM, row-aligned with the index (value + presence
    bit per row); predicates on it become mask handles without touching the host."""

    OPS = {None: 0, "$ne": 1, "$gt": 2, "$gte": 3, "$lt": 4, "$lte": 5}

    def __init__(self, engine: "FlatIPEngine"):
        self._engine = engine
        self._h = ctypes.c_void_p()
        self.rows = 0
        N.check(N.lib().mvdb_column_create(engine.handle, ctypes.byref(self._h)))

    def append(self, values: np.ndarray, present: np.ndarray) -> None:
        values = np.ascontiguousarray(values, dtype=np.float64)
        present = np.ascontiguousarray(present, dtype=np.uint8)
        assert values.shape == present.shape
        N.check(N.lib().mvdb_column_append(self._h, values.ctypes.data, present.ctypes.data, values.shape[0]))
        self.rows += int(values.shape[0])

    def predicate(self, op, operand: float) -> MaskHandle:
        h = ctypes.c_void_p()
        N.check(N.lib().mvdb_mask_from_predicate(self._engine.handle, self._h, self.OPS[op], float(operand), ctypes.byref(h)))
        return MaskHandle(self._engine, _raw=h, _rows=self.rows)

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            N.lib().mvdb_column_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class FlatIPEngine:
    def __init__(self, d: int, device: int = 0, capacity_hint: int = 0):
        self._h = ctypes.c_void_p()
        self.d = int(d)
        self.device = int(device)
        N.check(N.lib().mvdb_index_create(self.d, self.device, int(capacity_hint), ctypes.byref(self._h)))

    # -- lifetime ----------------------------------------------------------
    def close(self) -> None:
        """Destroy the index.  Mask handles, columns and workspaces made from it become inert (their own
        close() stays legal).  An engine that belongs to a ShardGroup cannot be closed before the group."""
        if getattr(self, "_h", None) is not None and self._h.value:
            N.check(N.lib().mvdb_index_destroy(self._h))
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    def reset(self) -> None:
        N.check(N.lib().mvdb_index_reset(self._h))

    def set_option(self, name: str, value: int) -> None:
        N.check(N.lib().mvdb_index_set_option(self._h, name.encode(), int(value)))

    # -- size ----------------------------------------------------------------
    def _counts(self) -> Tuple[int, int]:
        nt, nl = ctypes.c_int64(), ctypes.c_int64()
        N.check(N.lib().mvdb_index_ntotal(self._h, ctypes.byref(nt), ctypes.byref(nl)))
        return nt.value, nl.value

    @property
    def ntotal(self) -> int:
        return self._counts()[0]

    @property
    def nlive(self) -> int:
        return self._counts()[1]

    # -- ingest ---------------------------------------------------------------
    def add(self, x, normalize: bool = False) -> int:
        """index.add(x) (VDB:46); normalize=True fuses faiss.normalize_L2 (VDB:45).
        Returns the row number of x[0]."""
        x = _as_f32_2d(x, self.d, "rows")
        first = ctypes.c_int64()
        N.check(N.lib().mvdb_index_add(self._h, x.ctypes.data, x.shape[0], int(bool(normalize)),
                                       ctypes.byref(first)))
        return first.value

    def add_device(self, ptr: int, n: int, normalize: bool = False) -> int:
        first = ctypes.c_int64()
        N.check(N.lib().mvdb_index_add_device(self._h, ctypes.c_void_p(ptr), int(n), int(bool(normalize)),
                                              ctypes.byref(first)))
        return first.value

    def add_synthetic(self, seed: int, row0: int, n: int, dist: int = 0, normalize: bool = True) -> int:
        first = ctypes.c_int64()
        N.check(N.lib().mvdb_index_add_synthetic(self._h, int(seed), int(row0), int(n), int(dist),
                                                 int(bool(normalize)), ctypes.byref(first)))
        return first.value

    # -- delete ---------------------------------------------------------------
    def remove_rows(self, rows) -> None:
        rows = np.ascontiguousarray(rows, dtype=np.int64).ravel()
        N.check(N.lib().mvdb_index_remove_rows(self._h, rows.ctypes.data, rows.shape[0]))

    def compact(self) -> int:
        nt = ctypes.c_int64()
        N.check(N.lib().mvdb_index_compact(self._h, ctypes.byref(nt)))
        return nt.value

    # -- read back --------------------------------------------------------------
    def reconstruct(self, row: int) -> np.ndarray:
        out = np.empty(self.d, dtype=np.float32)
        N.check(N.lib().mvdb_index_reconstruct(self._h, int(row), out.ctypes.data))
        return out

    def reconstruct_n(self, row0: int, n: int) -> np.ndarray:
        out = np.empty((int(n), self.d), dtype=np.float32)
        N.check(N.lib().mvdb_index_reconstruct_n(self._h, int(row0), int(n), out.ctypes.data))
        return out

    def device_view(self):
        """(matrix device pointer, leading dimension in floats, live-mask device pointer)."""
        m, l = ctypes.c_void_p(), ctypes.c_void_p()
        ld = ctypes.c_int64()
        N.check(N.lib().mvdb_index_device_view(self._h, ctypes.byref(m), ctypes.byref(ld), ctypes.byref(l)))
        return m.value, ld.value, l.value

    # -- search ---------------------------------------------------------------
    def search(self, q, k: int, mask: Optional[np.ndarray] = None, mask_rows: Optional[int] = None,
               normalize: bool = False):
        """index.search(q, k) (VDB:497).  `mask` is either a bool[n] array of
        admissible rows or already-packed bytes (then pass mask_rows).
        Returns (D float32[nq,k], I int64[nq,k]) with faiss padding."""
        q = _as_f32_2d(q, self.d, "queries")
        k = int(k)
        if k <= 0:
            raise ValueError("k must be positive")
        nq = q.shape[0]
        D = np.empty((nq, k), dtype=np.float32)
        I = np.empty((nq, k), dtype=np.int64)
        if isinstance(mask, MaskHandle):
            N.check(N.lib().mvdb_index_search_with_mask(self._h, q.ctypes.data, nq, k, mask._h, int(bool(normalize)),
                                                        D.ctypes.data, I.ctypes.data))
            return D, I
        mptr, mrows = None, 0
        if mask is not None:
            mask = np.asarray(mask)
            if mask.dtype == np.bool_:
                mrows = mask.shape[0]
                mask = pack_mask(mask)
            else:
                mask = np.ascontiguousarray(mask, dtype=np.uint8)
                mrows = int(mask_rows) if mask_rows is not None else mask.shape[0] * 8
                if mrows > mask.shape[0] * 8:
                    raise ValueError("mask_rows exceeds the mask length")
            mptr = mask.ctypes.data
        N.check(N.lib().mvdb_index_search(self._h, q.ctypes.data, nq, k, mptr, mrows,
                                          int(bool(normalize)), D.ctypes.data, I.ctypes.data))
        return D, I

    def mask_handle(self, admissible) -> MaskHandle:
        """Upload a bool[n] filter once; reuse it across searches."""
        return MaskHandle(self, admissible)

    def mask_filled(self, rows: int) -> MaskHandle:
        """Device-resident mask admitting rows [0, rows)."""
        h = ctypes.c_void_p()
        N.check(N.lib().mvdb_mask_create_filled(self._h, int(rows), ctypes.byref(h)))
        return MaskHandle(self, _raw=h, _rows=rows)

    def column(self) -> DeviceColumn:
        return DeviceColumn(self)

    def debug_gemm_scores(self, q) -> np.ndarray:
        """Raw bf16 tensor-core scores [nq, ntotal] (test hook for the tcgen05 GEMM)."""
        q = _as_f32_2d(q, self.d, "queries")
        out = np.empty((q.shape[0], self.ntotal), dtype=np.float32)
        N.check(N.lib().mvdb_debug_gemm_scores(self._h, q.ctypes.data, q.shape[0], out.ctypes.data))
        return out

    # -- device-buffer flavour (sharded path, bench) -----------------------------
    def workspace(self) -> "Workspace":
        return Workspace(self)

    def search_device(self, ws: "Workspace", q_ptr: int, nq: int, k: int, D_ptr: int, I_ptr: int,
                      mask_ptr: int = 0, mask_rows: int = 0, normalize: bool = False,
                      label_offset: int = 0, stream: int = 0) -> None:
        N.check(N.lib().mvdb_index_search_device(
            self._h, ws._h, ctypes.c_void_p(q_ptr), int(nq), int(k),
            ctypes.c_void_p(mask_ptr) if mask_ptr else None, int(mask_rows), int(bool(normalize)),
            int(label_offset), ctypes.c_void_p(D_ptr), ctypes.c_void_p(I_ptr),
            ctypes.c_void_p(stream) if stream else None))


class ShardGroup:
    """Several FlatIPEngines (one per GPU of one box) searched as ONE index by one process
    (mvdb_group_*): every device scans its shard concurrently, the scans exchange their top-k over
    NVLink inside the kernel and the merged list comes back from shard 0.  This is what the
    reference's single in-memory index (ref sharded_vector_database.py:79-84) becomes when
    `ShardedVectorDatabase(devices=[...])` spreads it over GPUs."""

    SHARD_SHIFT = 40
    K_MAX = 128

    def __init__(self, engines):
        self.engines = list(engines)
        self._h = ctypes.c_void_p()
        arr = (ctypes.c_void_p * len(self.engines))(*[e.handle for e in self.engines])
        N.check(N.lib().mvdb_group_create(arr, len(self.engines), ctypes.byref(self._h)))

    def set_option(self, name: str, value: int) -> None:
        N.check(N.lib().mvdb_group_set_option(self._h, name.encode(), int(value)))

    def search(self, q, k: int, masks=None, normalize: bool = False):
        """masks: None or one entry per shard -- None, a MaskHandle of that shard's engine, or a bool[n_shard]
        array.  Returns (D [nq,k], shard [nq,k], row [nq,k]); unfilled slots have shard = row = -1."""
        d = self.engines[0].d
        q = _as_f32_2d(q, d, "queries")
        k = int(k)
        if k <= 0:
            raise ValueError("k must be positive")
        nq, ns = q.shape[0], len(self.engines)
        D = np.empty((nq, k), dtype=np.float32)
        I = np.empty((nq, k), dtype=np.int64)
        hp = mp = rp = None
        keep = []
        if masks is not None and any(m is not None for m in masks):
            handles = (ctypes.c_void_p * ns)()
            hosts = (ctypes.c_void_p * ns)()
            rows = (ctypes.c_uint64 * ns)()
            for i, m in enumerate(masks):
                if m is None:
                    continue   # both pointers NULL: this shard is searched unfiltered
                if isinstance(m, MaskHandle):
                    handles[i] = m._h.value
                    keep.append(m)
                else:
                    m = np.asarray(m, dtype=bool)
                    packed = pack_mask(m)
                    keep.append(packed)
                    hosts[i] = packed.ctypes.data if packed.size else None
                    rows[i] = m.shape[0]
                    if not packed.size:   # an empty mask must still read as "nothing admissible"
                        z = np.zeros(1, dtype=np.uint8)
                        keep.append(z)
                        hosts[i] = z.ctypes.data
            hp, mp, rp = handles, hosts, rows
        N.check(N.lib().mvdb_group_search(self._h, q.ctypes.data, nq, k, hp, mp, rp, int(bool(normalize)),
                                          D.ctypes.data, I.ctypes.data))
        shard = np.where(I >= 0, I >> self.SHARD_SHIFT, -1)
        row = np.where(I >= 0, I & ((1 << self.SHARD_SHIFT) - 1), -1)
        return D, shard, row

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            N.lib().mvdb_group_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


FlatIPEngine.group_class = ShardGroup   # what _store ties several partitions' engines together with


class Workspace:
    def __init__(self, engine: FlatIPEngine):
        self._h = ctypes.c_void_p()
        self._engine = engine  # keep the index alive
        N.check(N.lib().mvdb_index_workspace_create(engine.handle, ctypes.byref(self._h)))

    def shadow_counters(self):
        """(candidates, survivors, overflowed) of the last int8 shadow search run on this workspace."""
        out = (ctypes.c_uint32 * 4)()
        N.check(N.lib().mvdb_debug_read_shadow_counters(self._h, out))
        return int(out[0]), int(out[1]), bool(out[2] & 1)

    def survivor_counters(self):
        """(length of the survivor list, overflowed) of the last survivor-tail scan (fp32, 32 < k <= 128) on this workspace."""
        out = (ctypes.c_uint32 * 4)()
        N.check(N.lib().mvdb_debug_read_shadow_counters(self._h, out))
        return int(out[3]), bool(out[2] & 2)

    def close(self):
        if self._h is not None and self._h.value:
            N.lib().mvdb_index_workspace_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def normalize_L2(x: np.ndarray, device: int = 0) -> None:
    """faiss.normalize_L2 (VDB:475): in place on a C-contiguous float32 [n, d] array."""
    if not (isinstance(x, np.ndarray) and x.dtype == np.float32 and x.ndim == 2 and x.flags.c_contiguous):
        raise TypeError("normalize_L2 needs a C-contiguous float32 [n, d] array")
    N.check(N.lib().mvdb_normalize_L2(x.ctypes.data, x.shape[0], x.shape[1], int(device)))


def merge_topk_device(device: int, D_parts_ptr: int, I_parts_ptr: int, nparts: int, nq: int, k: int,
                      D_out_ptr: int, I_out_ptr: int, stream: int = 0) -> None:
    N.check(N.lib().mvdb_merge_topk_device(int(device), ctypes.c_void_p(D_parts_ptr), ctypes.c_void_p(I_parts_ptr),
                                           int(nparts), int(nq), int(k), ctypes.c_void_p(D_out_ptr),
                                           ctypes.c_void_p(I_out_ptr),
                                           ctypes.c_void_p(stream) if stream else None))
