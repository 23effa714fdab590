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


class FakeGroup:
    """CPU stand-in for ShardGroup (mvdb_group_*): every shard searched by the oracle, lists merged by
    score descending with exact ties ordered by (shard, row) -- the order the fused kernel produces."""
    K_MAX = 128
    searches = 0

    def __init__(self, engines):
        self.engines = list(engines)

    def search(self, q, k, masks=None, normalize=False):
        FakeGroup.searches += 1
        q = np.ascontiguousarray(q, dtype=np.float32).reshape(-1, self.engines[0].d)
        nq = q.shape[0]
        D = np.full((nq, k), O.FLT_LOWEST, dtype=np.float32)
        S = np.full((nq, k), -1, dtype=np.int64)
        R = np.full((nq, k), -1, dtype=np.int64)
        per = []
        for s, e in enumerate(self.engines):
            m = None if masks is None else masks[s]
            if e.ntotal == 0 or (m is not None and len(m) == 0):
                continue
            per.append((s, e.search(q, k, mask=m, normalize=normalize)))
        for i in range(nq):
            c = [(-float(d), s, int(r)) for s, (Dd, Ii) in per for d, r in zip(Dd[i], Ii[i]) if r >= 0]
            c.sort()
            for j, (nd, s, r) in enumerate(c[:k]):
                D[i, j], S[i, j], R[i, j] = -nd, s, r
        return D, S, R

    def close(self):
        pass


FakeEngine.group_class = FakeGroup
