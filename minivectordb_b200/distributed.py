"""Row-sharded search across the GPUs of one box: one process per GPU.

The reference's `ShardedVectorDatabase` shards FILES, not the search (ref
minivectordb/sharded_vector_database.py:79-84, 134-154: one in-memory index
over all rows).  BASELINE.json's north_star maps it onto the hardware instead:
each rank keeps a contiguous block of rows resident in its GPU's HBM, every
query is scanned by all ranks at once (no data-path collective: rows are
independent units), and the only exchange is k (score, label) pairs per rank
-- an all-gather over NVLink followed by a k-way merge that every rank runs
redundantly (so every rank holds the answer; 8 x k x 12 B at k=10 is 960 B,
latency-bound).

Two exchange transports:
* "fused" (default when CUDA IPC works): the scan kernel's last CTA stores the
  rank's k best keys straight into every peer's receive buffer over NVLink
  (peer-mapped memory), publishes a sequence flag, waits for the peers' flags
  and merges -- scan + exchange + merge in ONE kernel launch, no NCCL call on
  the query path (csrc/scan.cuh, xchg_*).
* "nccl": scan kernel, two ncclAllGather (scores, labels), merge kernel.

Labels are global row numbers: rank r adds `offset_r` (rows held by lower
ranks) to its local row numbers inside the scan epilogue, and the merge breaks
exact score ties by (rank, local order) = ascending global row, the same rule
the single-GPU scan uses.
"""
from __future__ import annotations

from typing import Callable, Optional

import numpy as np
import torch
import torch.distributed as dist

import ctypes

from . import _native as N
from .engine import FlatIPEngine, merge_topk_device


class ExchangeTimeout(RuntimeError):
    """A rank's scan gave up waiting for a peer's top-k (the peer died or never issued the search)."""


class RowShardedIndex:
    FUSED_K_MAX = 128
    FUSED_NQ_MAX = 8
    STABLE_SHIFT = 40      # numbering="stable": label = rank << 40 | row of that rank's shard

    def __init__(self, d: int, device: Optional[int] = None, group=None,
                 engine_factory: Callable[..., FlatIPEngine] = None, host_merge=None, exchange: str = "auto",
                 numbering: str = "contiguous"):
        """`engine_factory` / `host_merge` exist for the CPU (gloo) tests of the
        exchange logic; the product path uses the CUDA engine and merge kernel.

        numbering="contiguous": labels are global row numbers, rank r holding rows [offset_r, offset_r + n_r)
        (a read-mostly index loaded in rank order).  numbering="stable": label = rank << 40 | local row, which
        never changes when other ranks grow -- the numbering `add_balanced` / `remove` need (ref
        sharded_vector_database.py:104-132 inserts into the first non-full shard, :206-241 deletes by id)."""
        if numbering not in ("contiguous", "stable"):
            raise ValueError("numbering must be 'contiguous' or 'stable'")
        self.numbering = numbering
        self.d = int(d)
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.on_gpu = engine_factory is None
        if self.on_gpu:
            self.device = torch.cuda.current_device() if device is None else int(device)
            self.engine = FlatIPEngine(self.d, device=self.device)
            self._ws = self.engine.workspace()
            self._tdev = torch.device("cuda", self.device)
        else:
            self.device = -1
            self.engine = engine_factory(self.d)
            self._tdev = torch.device("cpu")
        self._host_merge = host_merge
        self.offset = 0        # global row number of this rank's row 0
        self.offsets = [0] * self.world
        self.ntotal_global = 0
        self._bufs = {}
        self._stage = {}
        self._xchg = None
        self.exchange = "none" if self.world == 1 else "nccl"
        if self.on_gpu and self.world > 1 and exchange in ("auto", "fused"):
            self._setup_fused(required=(exchange == "fused"))

    def _setup_fused(self, required: bool) -> None:
        """Create the peer-mapped exchange buffers and swap CUDA-IPC handles."""
        L = N.lib()
        h = ctypes.c_void_p()
        ok, handle = 1, bytes(64)
        try:
            N.check(L.mvdb_exchange_create(self.device, self.rank, self.world, self.FUSED_K_MAX, self.FUSED_NQ_MAX,
                                           ctypes.byref(h)))
            buf = ctypes.create_string_buffer(64)
            N.check(L.mvdb_exchange_ipc_handle(h, buf))
            handle = buf.raw
        except N.MvdbError:
            ok = 0
        gathered = [None] * self.world
        dist.all_gather_object(gathered, (ok, handle), group=self.group)
        if all(g[0] for g in gathered):
            blob = b"".join(g[1] for g in gathered)
            offs = (ctypes.c_int64 * self.world)(*self.offsets)
            try:
                N.check(L.mvdb_exchange_connect(h, blob, offs))
            except N.MvdbError:
                ok = 0
        flags = [None] * self.world
        dist.all_gather_object(flags, ok, group=self.group)
        if all(flags):
            self._xchg = h
            self.exchange = "fused"
        else:
            if h.value:
                L.mvdb_exchange_destroy(h)
            if required:
                raise RuntimeError("fused exchange unavailable (CUDA IPC / peer access failed on some rank)")

    # -- ingest (collective: every rank calls it, possibly with zero rows) ---------
    def add(self, x=None, normalize: bool = True, synthetic=None) -> None:
        """Append this rank's block.  Ranks hold contiguous global row ranges in
        rank order, so blocks must be added in one collective step per batch;
        offsets are re-derived from an all-gather of the per-rank row counts."""
        if synthetic is not None:
            seed, row0, n, dist_kind = synthetic
            self.engine.add_synthetic(seed, row0, n, dist_kind, normalize)
        elif x is not None and len(x):
            self.engine.add(np.ascontiguousarray(x, dtype=np.float32), normalize=normalize)
        self._sync_offsets()

    def _sync_offsets(self) -> None:
        counts = torch.zeros(self.world, dtype=torch.int64, device=self._tdev)
        mine = torch.tensor([self.engine.ntotal], dtype=torch.int64, device=self._tdev)
        if self.world > 1:
            dist.all_gather_into_tensor(counts, mine, group=self.group)
        else:
            counts.copy_(mine)
        counts = counts.cpu().tolist()
        self.counts = [int(c) for c in counts]
        if self.numbering == "stable":
            self.offsets = [r << self.STABLE_SHIFT for r in range(self.world)]
        else:
            self.offsets = [int(sum(counts[:r])) for r in range(self.world)]
        self.offset = self.offsets[self.rank]
        self.ntotal_global = int(sum(counts))
        if self._xchg is not None:
            offs = (ctypes.c_int64 * self.world)(*self.offsets)
            N.check(N.lib().mvdb_exchange_set_offsets(self._xchg, offs))

    def add_balanced(self, x, normalize: bool = True) -> np.ndarray:
        """Collective insert of the SAME block `x` on every rank: the rows go, one by one, to the rank
        that holds the fewest rows (lowest rank on a draw) -- the multi-GPU reading of the reference's
        "first shard with room" (ref sharded_vector_database.py:98-132).  Needs numbering="stable".
        Returns the int64 label of every row of `x` (identical on every rank)."""
        if self.numbering != "stable":
            raise ValueError("add_balanced needs numbering='stable' (labels must survive other ranks' growth)")
        x = np.ascontiguousarray(x, dtype=np.float32).reshape(-1, self.d)
        if not hasattr(self, "counts"):
            self._sync_offsets()
        counts = list(self.counts)
        owner = np.empty(x.shape[0], dtype=np.int64)
        labels = np.empty(x.shape[0], dtype=np.int64)
        for i in range(x.shape[0]):   # every rank replays the same deterministic assignment
            r = min(range(self.world), key=lambda j: (counts[j], j))
            owner[i] = r
            labels[i] = (r << self.STABLE_SHIFT) | counts[r]
            counts[r] += 1
        mine = x[owner == self.rank]
        if mine.shape[0]:
            first = self.engine.add(mine, normalize=normalize)
            assert first == self.counts[self.rank]
        self._sync_offsets()
        assert self.counts == counts, "ranks disagree on the assignment (different blocks passed to add_balanced?)"
        return labels

    def remove(self, labels) -> None:
        """Collective delete by label (every rank passes the same labels; the owner tombstones its rows).
        Unknown or already deleted labels raise on the owning rank, as the reference raises for unknown
        ids (ref sharded_vector_database.py:213-217)."""
        labels = np.asarray(labels, dtype=np.int64).reshape(-1)
        if self.numbering == "stable":
            mine = labels[(labels >> self.STABLE_SHIFT) == self.rank] & ((1 << self.STABLE_SHIFT) - 1)
        else:
            hi = self.offset + self.engine.ntotal
            mine = labels[(labels >= self.offset) & (labels < hi)] - self.offset
        if mine.shape[0]:
            self.engine.remove_rows(mine)

    def set_exchange_timeout(self, ms: int) -> None:
        """How long a scan waits for a peer's top-k before it gives up (default 2000 ms)."""
        if self._xchg is not None:
            N.check(N.lib().mvdb_exchange_set_option(self._xchg, b"timeout_ms", int(ms)))

    def check_exchange(self) -> None:
        """Raise ExchangeTimeout if any search since the last check gave up on a peer.  Costs one read
        of pinned host memory; valid after a synchronise that covers the searches."""
        if self.exchange_timed_out():
            raise ExchangeTimeout(f"rank {self.rank}: a peer did not deliver its top-k within the exchange timeout")

    # -- search ---------------------------------------------------------------------
    def _buffers(self, nq: int, k: int):
        key = (nq, k)
        if key not in self._bufs:
            t = dict(
                D_loc=torch.empty((nq, k), dtype=torch.float32, device=self._tdev),
                I_loc=torch.empty((nq, k), dtype=torch.int64, device=self._tdev),
                D_parts=torch.empty((self.world, nq, k), dtype=torch.float32, device=self._tdev),
                I_parts=torch.empty((self.world, nq, k), dtype=torch.int64, device=self._tdev),
            )
            # results live in ONE buffer ([labels | distances]) so that a host caller needs one D2H transfer
            on = nq * k
            t["out"] = torch.empty(on + (on + 1) // 2, dtype=torch.int64, device=self._tdev)
            t["I_out"] = t["out"][:on].view(nq, k)
            t["D_out"] = t["out"][on:].view(torch.float32)[:on].view(nq, k)
            if self.on_gpu:
                t["out_pin"] = torch.empty_like(t["out"], device="cpu").pin_memory()
            self._bufs[key] = t
        return self._bufs[key]

    def search_device(self, q_dev: torch.Tensor, k: int, mask_dev: Optional[torch.Tensor] = None,
                      mask_rows: int = 0, normalize: bool = False):
        """Enqueue scan -> all-gather -> merge on the current CUDA stream.
        q_dev: float32 [nq, d] on this rank's GPU (same query on every rank).
        Returns device tensors (D [nq,k], I [nq,k] global rows), valid after the
        stream is synchronised."""
        nq = q_dev.shape[0]
        b = self._buffers(nq, k)
        st = torch.cuda.current_stream().cuda_stream
        if self._xchg is not None and k <= self.FUSED_K_MAX:
            N.check(N.lib().mvdb_index_search_exchange(
                self.engine.handle, self._ws._h, self._xchg, ctypes.c_void_p(q_dev.data_ptr()), nq, int(k),
                ctypes.c_void_p(mask_dev.data_ptr()) if mask_dev is not None else None, int(mask_rows),
                int(bool(normalize)), ctypes.c_void_p(b["D_out"].data_ptr()), ctypes.c_void_p(b["I_out"].data_ptr()),
                ctypes.c_void_p(st) if st else None))
            return b["D_out"], b["I_out"]
        # a lone rank writes straight into the result buffer (one D2H for host callers, see search_packed)
        D_loc, I_loc = (b["D_out"], b["I_out"]) if self.world == 1 else (b["D_loc"], b["I_loc"])
        self.engine.search_device(self._ws, q_dev.data_ptr(), nq, k, D_loc.data_ptr(), I_loc.data_ptr(),
                                  mask_ptr=mask_dev.data_ptr() if mask_dev is not None else 0,
                                  mask_rows=mask_rows, normalize=normalize, label_offset=self.offset, stream=st)
        if self.world == 1:
            return D_loc, I_loc
        dist.all_gather_into_tensor(b["D_parts"].view(-1), b["D_loc"].view(-1), group=self.group)
        dist.all_gather_into_tensor(b["I_parts"].view(-1), b["I_loc"].view(-1), group=self.group)
        merge_topk_device(self.device, b["D_parts"].data_ptr(), b["I_parts"].data_ptr(), self.world, nq, k,
                          b["D_out"].data_ptr(), b["I_out"].data_ptr(), st)
        return b["D_out"], b["I_out"]

    def search(self, q, k: int, mask_local: Optional[np.ndarray] = None, normalize: bool = False):
        """Host-array convenience: same query on every rank; `mask_local` is this
        rank's slice (bool[n_local]) of the admissible-row mask.  Returns numpy
        (D, I) with global row numbers, identical on every rank."""
        q = np.ascontiguousarray(q, dtype=np.float32).reshape(-1, self.d)
        nq = q.shape[0]
        if self.on_gpu:
            qd = torch.from_numpy(q).to(self._tdev)
            md, mrows = None, 0
            if mask_local is not None:
                mrows = int(mask_local.shape[0])
                packed = np.packbits(np.asarray(mask_local, dtype=bool), bitorder="little")
                words = np.zeros((mrows + 31) // 32 * 4, dtype=np.uint8)
                words[:packed.size] = packed
                md = torch.from_numpy(words.view(np.int32)).to(self._tdev)
            D, I = self.search_device(qd, k, md, mrows, normalize)
            torch.cuda.current_stream().synchronize()
            self.check_exchange()
            return D.cpu().numpy(), I.cpu().numpy()
        # CPU test path: injected engine + injected merge, gloo collectives
        D, I = self.engine.search(q, k, mask=mask_local, normalize=normalize)
        I = np.where(I >= 0, I + self.offset, -1)
        if self.world == 1:
            return D, I
        b = self._buffers(nq, k)
        dist.all_gather_into_tensor(b["D_parts"].view(-1), torch.from_numpy(D).contiguous().view(-1), group=self.group)
        dist.all_gather_into_tensor(b["I_parts"].view(-1), torch.from_numpy(I).contiguous().view(-1), group=self.group)
        return self._host_merge(b["D_parts"].numpy(), b["I_parts"].numpy(), k)

    def search_packed(self, q, k: int, mask_words: Optional[np.ndarray] = None, mask_rows: int = 0,
                      normalize: bool = False):
        """Host-buffer search with the fewest transfers: the packed filter of THIS rank's shard
        (uint8/int32 little-endian bit words, bit r = row r admissible, as `pack_mask` + zero padding to
        whole 32-bit words) and the query travel H2D as one pinned transfer, (labels, distances) come
        back as one.  Same query on every rank.  Returns numpy (D, I) with global row numbers."""
        q = np.ascontiguousarray(q, dtype=np.float32).reshape(-1, self.d)
        nq = q.shape[0]
        if not self.on_gpu:
            # CPU test path (gloo + injected engine): same contract, the packed words are unpacked again
            loc = None
            if mask_words is not None:
                bits = np.unpackbits(np.ascontiguousarray(mask_words).view(np.uint8), bitorder="little")
                loc = bits[:int(mask_rows)].astype(bool)
            return self.search(q, k, mask_local=loc, normalize=normalize)
        nw = 0
        if mask_words is not None:
            mw = np.ascontiguousarray(mask_words).view(np.int32).reshape(-1)
            nw = mw.shape[0]
        q_at = (nw + 31) // 32 * 32                      # query starts on a 128-byte boundary
        need = q_at + nq * self.d
        st = self._stage.get(need)
        if st is None:
            st = (torch.empty(need, dtype=torch.int32).pin_memory(), torch.empty(need, dtype=torch.int32, device=self._tdev))
            self._stage[need] = st
        pin, dev = st
        ph = pin.numpy()
        if nw:
            ph[:nw] = mw
        ph[q_at:need] = q.view(np.int32).reshape(-1)
        dev.copy_(pin, non_blocking=True)
        qd = dev[q_at:need].view(torch.float32).view(nq, self.d)
        D, I = self.search_device(qd, k, dev[:nw] if nw else None, mask_rows if nw else 0, normalize)
        b = self._buffers(nq, k)
        b["out_pin"].copy_(b["out"], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        self.check_exchange()
        on = nq * k
        host = b["out_pin"].numpy()
        return host[on:].view(np.float32)[:on].reshape(nq, k).copy(), host[:on].reshape(nq, k).copy()

    def exchange_timed_out(self) -> bool:
        if self._xchg is None:
            return False
        v = ctypes.c_int(0)
        N.check(N.lib().mvdb_exchange_status(self._xchg, ctypes.byref(v)))
        return bool(v.value)

    def close(self):
        if self._xchg is not None:
            torch.cuda.synchronize()
            if self.world > 1:
                dist.barrier(group=self.group)  # nobody unmaps while a peer may still write
            N.lib().mvdb_exchange_destroy(self._xchg)
            self._xchg = None
        if self.on_gpu:
            self._ws.close()
        self.engine.close()
