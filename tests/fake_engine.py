"""Oracle-backed stand-in for FlatIPEngine, for CPU-only tests of the HOST
logic (filters, id maps, staging, tombstones, compaction bookkeeping).  It is
test infrastructure: the product never falls back to it."""
import numpy as np

from oracle import oracle as O


class FakeEngine:
    instances = 0

    def __init__(self, d, device=0, capacity_hint=0):
        self.d, self.device = int(d), device
        self.x = np.zeros((0, self.d), dtype=np.float32)
        self.live = np.zeros(0, dtype=bool)
        FakeEngine.instances += 1

    ntotal = property(lambda self: self.x.shape[0])
    nlive = property(lambda self: int(self.live.sum()))

    def add(self, x, normalize=False):
        x = np.ascontiguousarray(x, dtype=np.float32).reshape(-1, self.d).copy()
        if normalize:
            O.normalize_L2(x)
        first = self.x.shape[0]
        self.x = np.vstack([self.x, x])
        self.live = np.concatenate([self.live, np.ones(x.shape[0], dtype=bool)])
        return first

    def remove_rows(self, rows):
        rows = np.asarray(rows, dtype=np.int64)
        assert self.live[rows].all() and len(set(rows.tolist())) == len(rows)
        self.live[rows] = False

    def compact(self):
        self.x = self.x[self.live]
        self.live = np.ones(self.x.shape[0], dtype=bool)
        return self.x.shape[0]

    def reconstruct(self, row):
        return self.x[int(row)].copy()

    def reconstruct_n(self, row0, n):
        return self.x[int(row0):int(row0) + int(n)].copy()

    def search(self, q, k, mask=None, mask_rows=None, normalize=False):
        q = np.ascontiguousarray(q, dtype=np.float32).reshape(-1, self.d).copy()
        if normalize:
            O.normalize_L2(q)
        adm = self.live.copy()
        if mask is not None:
            m = np.asarray(mask)
            if m.dtype != np.bool_:
                m = np.unpackbits(m.astype(np.uint8), bitorder="little")[:mask_rows].astype(bool)
            full = np.zeros(self.x.shape[0], dtype=bool)
            full[:m.shape[0]] = m[:self.x.shape[0]]
            adm &= full
        return O.search_masked(self.x, adm, q, int(k))

    def mask_handle(self, admissible):
        return np.asarray(admissible, dtype=bool).copy()

    def close(self):
        pass
