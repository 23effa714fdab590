"""Shared host-side core of the two drop-in database classes.

What the reference keeps as a numpy matrix + a faiss index + id dicts that are
all rebuilt or copied on every mutation (ref minivectordb/vector_database.py:
42-47, 57-155), this keeps as:

* one `FlatIPEngine` (HBM-resident matrix + tombstones) per partition -- one
  partition for `VectorDatabase`, one per device for `ShardedVectorDatabase`;
* append-only host arrays indexed by a global insertion id ("gid"): user id,
  metadata dict, live flag, (partition, slot);
* a columnar `FilterIndex` over gids that turns the mongo-like filters into
  the admissible-row bitmask the scan consumes.

Row numbers the user can observe (`id_map`, `inverse_id_map`, `metadata`
order, `embeddings[i]`) are the reference's dense LOGICAL rows: the rank of a
row among the live rows in insertion order.  Deleting therefore "renumbers"
exactly as the reference does (VDB:138-152) without touching the matrix;
physical compaction happens lazily and preserves order, so exact-tie order
(ascending row) is the same before and after.

Inserted rows are staged on the host, raw, and flushed to the GPU (normalised
by the ingest kernel) at the next search -- the analogue of the reference's
lazy `_build_index` (VDB:477-479), and the reason `get_vector` returns raw
values before the first search and normalised values after it, as the
reference does (VDB:45, 49-55).
"""
from __future__ import annotations

import threading
from collections import OrderedDict, defaultdict
from typing import List, Optional

import numpy as np

from . import rerank as _rerank
from .engine import FlatIPEngine
from .filters import FilterIndex


class _Partition:
    """One device-resident row shard and the host rows waiting to join it."""

    def __init__(self, device: int):
        self.device = device
        self.engine: Optional[FlatIPEngine] = None
        self.gids: List[int] = []        # slot -> gid, for flushed AND pending slots
        self.flushed = 0                 # slots [0, flushed) live on the device
        self.pending: List[np.ndarray] = []          # staged rows, one entry per slot (get_vector reads them)
        self.pending_blocks: List[np.ndarray] = []   # the same rows as 2-D blocks in arrival order (what a flush ships)
        self.dead_unflushed: List[int] = []  # slots deleted on the host, not yet tombstoned on the device
        self.n_dead = 0                  # tombstones currently held by the engine (+ unflushed ones)
        self._gids_np = None

    def gids_np(self) -> np.ndarray:
        if self._gids_np is None or self._gids_np.shape[0] != len(self.gids):
            self._gids_np = np.asarray(self.gids, dtype=np.int64)
        return self._gids_np


class EmbeddingsView:
    """Read-only stand-in for the reference's `self.embeddings` ndarray
    (VDB:12): len(), .shape, row indexing and np.asarray() in LOGICAL row order."""

    def __init__(self, store: "GpuStore"):
        self._s = store

    def __len__(self):
        return self._s._n_live

    @property
    def shape(self):
        return (self._s._n_live, self._s.embedding_size)

    @property
    def dtype(self):
        return np.dtype(np.float32)

    def __array__(self, dtype=None, copy=None):
        a = self._s._materialize()
        return a if dtype is None else a.astype(dtype)

    def __getitem__(self, idx):
        if isinstance(idx, (int, np.integer)):
            with self._s.lock:
                gids = self._s._live_gids()
                return self._s._row_of_gid(int(gids[idx]))
        return self._s._materialize()[idx]

    def __iter__(self):
        return iter(self._s._materialize())


class GpuStore:
    MASK_CACHE_ENTRIES = 16      # device-resident filter masks kept per database
    PARALLEL_PARTS_BYTES = 64 << 20   # partitions are scanned concurrently once the database is at least this big
    COMPACT_MIN_DEAD = 4096      # do not bother compacting below this many tombstones
    COMPACT_DEAD_FRACTION = 0.25

    def __init__(self, devices=None, scan_shadow=False):
        self.embedding_size = None
        self._scan_shadow = bool(scan_shadow)   # opt-in int8 shadow scan for single queries (same answers, ~1/4 of the bytes)
        self.lock = threading.Lock()
        self.inverted_index = defaultdict(set)   # metadata key -> set of unique ids (VDB:16, 78-79)
        self._devices = list(devices) if devices else [0]
        self._parts = [_Partition(dev) for dev in self._devices]
        self._g_uid: List = []
        self._g_meta: List[dict] = []
        self._g_part: List[int] = []
        self._g_slot: List[int] = []
        self._g_live = np.zeros(1024, dtype=bool)
        self._g_n = 0
        self._n_live = 0
        self._uid_gid = {}
        self._filters = FilterIndex()
        self._ever_stored = False
        self._active_searches = 0
        self._views = None               # cached (id_map, inverse_id_map, metadata, unique_ids)
        self._version = 0                # bumped by every mutation: invalidates cached filter masks
        self._pool = None                 # worker threads that scan several partitions (GPUs) at once
        self._mask_cache = OrderedDict()  # filter repr -> (version, count, per-partition device-resident masks)
        self._mask_seen = OrderedDict()   # filter repr -> version at which it was last evaluated
        self._embeddings_changed = False  # kept for API parity (VDB:18); True while rows wait on the host
        self._group = None                # several partitions: the engines searched as ONE index (fused NVLink exchange)

    # ------------------------------------------------------------------ views
    def _live_gids(self) -> np.ndarray:
        return np.flatnonzero(self._g_live[:self._g_n])

    def _build_views(self):
        if self._views is None:
            gids = self._live_gids().tolist()
            uids = [self._g_uid[g] for g in gids]
            self._views = ({i: u for i, u in enumerate(uids)}, {u: i for i, u in enumerate(uids)},
                           [self._g_meta[g] for g in gids], uids)
        return self._views

    @property
    def embeddings(self):
        return EmbeddingsView(self) if self._ever_stored else None

    def _row_of_gid(self, gid: int) -> np.ndarray:
        part = self._parts[self._g_part[gid]]
        slot = self._g_slot[gid]
        if slot >= part.flushed:
            return part.pending[slot - part.flushed]
        return part.engine.reconstruct(slot)

    def _materialize(self) -> np.ndarray:
        """All live rows, logical order, float32 [n_live, d] (what the reference's
        `self.embeddings` holds: normalised rows once flushed, raw while pending)."""
        with self.lock:
            return self._materialize_locked()

    def _materialize_locked(self) -> np.ndarray:
        """`_materialize` for callers that already hold the lock (persist_to_disk builds the
        matrix and the id views under ONE lock, as the reference does, VDB:538-548)."""
        d = self.embedding_size or 0
        gids = self._live_gids()
        out = np.empty((gids.shape[0], d), dtype=np.float32)
        if gids.shape[0] == 0:
            return out
        gpart = np.asarray(self._g_part, dtype=np.int64)[gids]
        gslot = np.asarray(self._g_slot, dtype=np.int64)[gids]
        for pi, part in enumerate(self._parts):
            sel = np.flatnonzero(gpart == pi)
            if sel.size == 0:
                continue
            slots = gslot[sel]
            on_dev = slots < part.flushed
            if on_dev.any():
                lo, hi = int(slots[on_dev].min()), int(slots[on_dev].max()) + 1
                block = part.engine.reconstruct_n(lo, hi - lo)
                out[sel[on_dev]] = block[slots[on_dev] - lo]
            for j in np.flatnonzero(~on_dev):
                out[sel[j]] = part.pending[int(slots[j]) - part.flushed]
        return out

    # --------------------------------------------------------------- mutation
    def _as_row(self, embedding) -> np.ndarray:
        row = np.array(embedding, dtype=np.float32)
        if row.ndim != 1:
            row = row.reshape(-1)
        if self.embedding_size is None:
            self.embedding_size = int(row.shape[0])
        elif row.shape[0] != self.embedding_size:
            # np.vstack raises the same type in the reference (VDB:72)
            raise ValueError(f"embedding has dimension {row.shape[0]}, database has {self.embedding_size}")
        return row

    def _as_rows(self, embeddings) -> np.ndarray:
        """A batch of embeddings as ONE float32 [m, d] block (same checks as `_as_row`, applied to
        every row; ragged or wrong-sized input raises ValueError before anything is stored)."""
        if isinstance(embeddings, np.ndarray) and embeddings.ndim == 2:
            # always a private copy: the block is staged until the next flush, and the reference copies
            # too (np.array + np.vstack, VDB:26, 107) -- a caller may reuse its batch buffer right away
            block = np.array(embeddings, dtype=np.float32, order='C')
        else:
            rows = [np.asarray(e, dtype=np.float32).reshape(-1) for e in embeddings]
            if not rows:
                return np.zeros((0, self.embedding_size or 0), dtype=np.float32)
            d0 = rows[0].shape[0] if self.embedding_size is None else self.embedding_size
            for r in rows:
                if r.shape[0] != d0:
                    raise ValueError(f"embedding has dimension {r.shape[0]}, database has {d0}")
            block = np.stack(rows)
        if block.shape[0] == 0:
            return block
        if self.embedding_size is None:
            self.embedding_size = int(block.shape[1])
        elif block.shape[1] != self.embedding_size:
            raise ValueError(f"embedding has dimension {block.shape[1]}, database has {self.embedding_size}")
        return block

    def _append_batch(self, uids, block: np.ndarray, metas) -> None:
        """`_append` for m rows at once (caller holds the lock and has validated uids / dimension):
        the bookkeeping lists grow by bulk extends and the only per-row Python left is the metadata
        walk that feeds the filter columns and the inverted index."""
        m = int(block.shape[0])
        if m == 0:
            return
        if len(self._parts) > 1:
            for uid, row, meta in zip(uids, block, metas):   # rows are balanced one by one across partitions
                self._append(uid, row, meta)
            return
        gid0 = self._g_n
        if gid0 + m > self._g_live.shape[0]:
            grown = np.zeros(max(self._g_live.shape[0] * 2, gid0 + m), dtype=bool)
            grown[:gid0] = self._g_live[:gid0]
            self._g_live = grown
        part = self._parts[0]
        slot0 = len(part.gids)
        gids = range(gid0, gid0 + m)
        uids = list(uids)
        metas = list(metas)
        self._g_uid.extend(uids)
        self._g_meta.extend(metas)
        self._g_part.extend([0] * m)
        self._g_slot.extend(range(slot0, slot0 + m))
        part.gids.extend(gids)
        part.pending.extend(block)          # one view per row: get_vector reads staged rows by slot
        part.pending_blocks.append(block)
        self._g_live[gid0:gid0 + m] = True
        self._g_n = gid0 + m
        self._n_live += m
        self._uid_gid.update(zip(uids, gids))
        columns, inv = self._filters.columns, self.inverted_index
        new_column = type(self._filters).new_column
        for gid, uid, meta in zip(gids, uids, metas):
            for key, value in meta.items():
                col = columns.get(key)
                if col is None:
                    col = columns[key] = new_column()
                col.rows.append(gid)
                col.vals.append(value)
                inv[key].add(uid)
        self._ever_stored = True
        self._embeddings_changed = True
        self._views = None
        self._version += 1

    def _append(self, uid, row: np.ndarray, metadata: dict) -> None:
        """Caller holds the lock and has validated uid / dimension."""
        gid = self._g_n
        if gid >= self._g_live.shape[0]:
            grown = np.zeros(self._g_live.shape[0] * 2, dtype=bool)
            grown[:gid] = self._g_live[:gid]
            self._g_live = grown
        pi = min(range(len(self._parts)), key=lambda i: len(self._parts[i].gids)) if len(self._parts) > 1 else 0
        part = self._parts[pi]
        self._g_uid.append(uid)
        self._g_meta.append(metadata)
        self._g_part.append(pi)
        self._g_slot.append(len(part.gids))
        part.gids.append(gid)
        part.pending.append(row)
        part.pending_blocks.append(row[None, :])
        self._g_live[gid] = True
        self._g_n = gid + 1
        self._n_live += 1
        self._uid_gid[uid] = gid
        self._filters.add_row(gid, metadata)
        for key in metadata:
            self.inverted_index[key].add(uid)
        self._ever_stored = True
        self._embeddings_changed = True
        self._views = None
        self._version += 1

    def _remove(self, uid) -> None:
        """Caller holds the lock and has checked that uid exists."""
        gid = self._uid_gid.pop(uid)
        self._g_live[gid] = False
        self._n_live -= 1
        part = self._parts[self._g_part[gid]]
        part.dead_unflushed.append(self._g_slot[gid])
        part.n_dead += 1
        for key in self._g_meta[gid]:
            ids = self.inverted_index.get(key)
            if ids is not None:
                ids.discard(uid)
                if not ids:
                    del self.inverted_index[key]
        self._embeddings_changed = True
        self._views = None
        self._version += 1

    def _new_engine(self, device):
        eng = FlatIPEngine(self.embedding_size, device=device)
        if self._scan_shadow and hasattr(eng, "set_option"):
            eng.set_option("scan_shadow", 1)
        return eng

    # ------------------------------------------------------------------ flush
    def _flush(self) -> None:
        """Move staged rows / tombstones to the GPU.  Caller holds the lock."""
        if len(self._parts) > 1 and self._group is None and any(p.pending for p in self._parts):
            # several GPUs: every partition gets its engine now and the engines are tied into one shard
            # group -- one host thread launches every device's scan and the scans exchange and merge
            # their top-k over NVLink inside the kernel (no thread pool, no Python merge)
            for part in self._parts:
                if part.engine is None:
                    part.engine = self._new_engine(part.device)
            group_cls = getattr(type(self._parts[0].engine), "group_class", None)
            if group_cls is not None:
                self._group = group_cls([p.engine for p in self._parts])
        for part in self._parts:
            if part.pending:
                if part.engine is None:
                    part.engine = self._new_engine(part.device)
                # a bulk store arrives as one block and is shipped as it is; np.vstack over a million
                # row views cost seconds here
                blocks = part.pending_blocks
                block = blocks[0] if len(blocks) == 1 else np.concatenate(blocks)
                part.engine.add(block, normalize=True)   # faiss.normalize_L2 + index.add (VDB:45-46)
                part.flushed = len(part.gids)
                part.pending = []
                part.pending_blocks = []
            if part.dead_unflushed:
                part.engine.remove_rows(np.asarray(part.dead_unflushed, dtype=np.int64))
                part.dead_unflushed = []
        self._embeddings_changed = False
        dead = self._g_n - self._n_live
        if (dead >= self.COMPACT_MIN_DEAD and dead >= self.COMPACT_DEAD_FRACTION * self._g_n
                and self._active_searches == 0):
            self._compact()

    def _compact(self) -> None:
        """Order-preserving squeeze of tombstones on the device and of the host
        arrays.  Caller holds the lock; no search is in flight."""
        keep = self._live_gids()
        remap = np.full(self._g_n, -1, dtype=np.int64)
        remap[keep] = np.arange(keep.shape[0])
        for part in self._parts:
            if part.engine is not None and part.n_dead:
                part.engine.compact()
            part.gids = [int(remap[g]) for g in part.gids if remap[g] >= 0]
            part.flushed = len(part.gids)
            part.n_dead = 0
            part._gids_np = None
        keep_l = keep.tolist()
        self._g_uid = [self._g_uid[g] for g in keep_l]
        self._g_meta = [self._g_meta[g] for g in keep_l]
        self._g_part = [self._g_part[g] for g in keep_l]
        self._g_n = len(keep_l)
        self._g_slot = [0] * self._g_n
        for part in self._parts:
            for slot, g in enumerate(part.gids):
                self._g_slot[g] = slot
        live = np.zeros(max(1024, 2 * self._g_n), dtype=bool)
        live[:self._g_n] = True
        self._g_live = live
        self._uid_gid = {u: g for g, u in enumerate(self._g_uid)}
        self._filters.clear()
        for g, meta in enumerate(self._g_meta):
            self._filters.add_row(g, meta)
        self._views = None
        self._version += 1
        self._mask_cache.clear()

    # ----------------------------------------------------------------- search
    def _search(self, embedding, metadata_filter, exclude_filter, or_filters, k, autocut):
        if not self._ever_stored:
            return [], [], []
        q = np.array(embedding, dtype=np.float32).reshape(1, -1)
        with self.lock:
            self._flush()
            n = self._g_n
            live = self._g_live[:n]
            jobs = None
            if metadata_filter or exclude_filter or or_filters:
                # A filter that repeats (same expression, no mutation since) reuses its
                # device-resident masks: no evaluation, no upload, and concurrent callers
                # can be coalesced into one tensor-core batch with per-query filters.
                try:
                    key = repr((metadata_filter, exclude_filter, or_filters))
                except Exception:  # noqa: BLE001 - unreprable operands: just do not cache
                    key = None
                hit = self._mask_cache.get(key) if key is not None else None
                if hit is not None and hit[0] == self._version:
                    self._mask_cache.move_to_end(key)
                    count, jobs = hit[1], hit[2]
                    adm = False  # placeholder: "filtered"
                elif (len(self._parts) == 1 and self._parts[0].engine is not None
                      and hasattr(self._parts[0].engine, "column") and self._parts[0].flushed == n):
                    # single GPU: the filter is evaluated ON the device (numeric predicates by a
                    # kernel over HBM-resident columns, the rest uploaded once) and combined there
                    part = self._parts[0]
                    handle = self._filters.admissible_device(part.engine, n, metadata_filter, exclude_filter, or_filters)
                    count = handle.count()
                    if count == 0:
                        return [], [], []
                    jobs = [(part, part.gids, handle)]
                    adm = False
                    if key is not None:
                        self._mask_cache[key] = (self._version, count, jobs)
                        while len(self._mask_cache) > self.MASK_CACHE_ENTRIES:
                            self._mask_cache.popitem(last=False)
                else:
                    adm = self._filters.admissible(live, metadata_filter, exclude_filter, or_filters)
            else:
                adm = None
            if jobs is None:
                count = self._n_live if adm is None else int(adm.sum())
            if count == 0:
                return [], [], []
            if jobs is None:
                if adm is not None and count == self._n_live:
                    adm = None  # every live row admissible: plain search (VDB:495-497)
                jobs = []
                for part in self._parts:
                    if part.engine is None or part.flushed == 0:
                        continue
                    if adm is None:
                        jobs.append((part, part.gids, None))
                    else:
                        pm = adm[part.gids_np()] if len(self._parts) > 1 or part.flushed != n else adm
                        # a filter seen for the first time travels as host bytes (no device
                        # allocation); when it comes back it is promoted to a resident handle
                        repeat = key is not None and self._mask_seen.get(key) == self._version
                        jobs.append((part, part.gids, part.engine.mask_handle(pm) if repeat else pm))
                if adm is not None and key is not None:
                    if self._mask_seen.get(key) == self._version:
                        self._mask_cache[key] = (self._version, count, jobs)
                        while len(self._mask_cache) > self.MASK_CACHE_ENTRIES:
                            self._mask_cache.popitem(last=False)  # handles are freed when their last user drops them
                    else:
                        self._mask_seen[key] = self._version
                        while len(self._mask_seen) > 4 * self.MASK_CACHE_ENTRIES:
                            self._mask_seen.popitem(last=False)
            g_uid, g_meta, g_live = self._g_uid, self._g_meta, self._g_live
            self._active_searches += 1
        try:
            search_k = min(int(k), count)  # VDB:489-492
            cands = []
            results = None
            if self._group is not None and search_k <= self._group.K_MAX:
                # ONE call: every device scans its partition at the same time, the per-device top-k are exchanged
                # over NVLink and merged by the scan kernels themselves; (shard, row) pairs come back merged
                masks = None
                if any(j[2] is not None for j in jobs):
                    masks = [np.zeros(0, dtype=bool)] * len(self._parts)   # partitions without a job hold nothing admissible
                    for part, gids, handle in jobs:
                        masks[self._parts.index(part)] = handle
                D, S, R = self._group.search(q, search_k, masks=masks, normalize=True)
                for dist, s_, r_ in zip(D[0], S[0], R[0]):
                    if s_ >= 0:
                        cands.append((dist, self._parts[int(s_)].gids[int(r_)]))
                cands.sort(key=lambda c: (-c[0], c[1]))  # exact ties: insertion order, as one index would return them
                results, jobs = [], []
            elif len(jobs) > 1 and self._n_live * (self.embedding_size or 1) * 4 >= self.PARALLEL_PARTS_BYTES:
                # partitions live on different GPUs: scan them at the same time (the C ABI releases the GIL);
                # below ~64 MB the thread hand-off costs more than the scans
                if self._pool is None:
                    from concurrent.futures import ThreadPoolExecutor
                    self._pool = ThreadPoolExecutor(max_workers=len(self._parts), thread_name_prefix="mvdb-part")
                results = list(self._pool.map(lambda j: j[0].engine.search(q, search_k, mask=j[2], normalize=True), jobs))
            else:
                results = [part.engine.search(q, search_k, mask=handle, normalize=True) for part, gids, handle in jobs]
            n_jobs = len(jobs)
            for (part, gids, handle), (D, I) in zip(jobs, results):
                for slot, dist in zip(I[0], D[0]):
                    if slot == -1:
                        continue  # VDB:500
                    cands.append((dist, gids[slot]))
            if n_jobs > 1:
                cands.sort(key=lambda c: (-c[0], c[1]))  # score desc, insertion order on exact ties
                cands = cands[:search_k]
            found = [(g_uid[g], dist, g_meta[g]) for dist, g in cands if g_live[g]]
        finally:
            with self.lock:
                self._active_searches -= 1
        ids, distances, metadatas = zip(*found) if found else ([], [], [])
        if autocut and len(distances) > 1:
            drop = set(_rerank.autocut_scores(distances))
            if drop:
                ids = [v for i, v in enumerate(ids) if i not in drop]
                distances = [v for i, v in enumerate(distances) if i not in drop]
                metadatas = [v for i, v in enumerate(metadatas) if i not in drop]
        return ids, distances, metadatas

    def __del__(self):
        # finalisers run in no particular order: release the native objects in THE order they need
        # (mask handles and columns, then the shard group, then the engines)
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    def close(self) -> None:
        if self._pool is not None:
            self._pool.shutdown(wait=True)
            self._pool = None
        # native mask / column handles point into their index: release them BEFORE the engines go
        for _, _, jobs in self._mask_cache.values():
            for job in jobs:
                if hasattr(job[2], "close"):
                    job[2].close()
        self._mask_cache.clear()
        self._mask_seen.clear()
        self._filters.close_device()
        if self._group is not None:
            self._group.close()
            self._group = None
        for part in self._parts:
            if part.engine is not None:
                part.engine.close()
                part.engine = None
