"""oracle/oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Python face of the CPU checker for MiniVectorDB's flat inner-product hot path.
PARITY UNPINNED (see the header of oracle/faiss_flat_ip.c): faiss-cpu, the
library the reference delegates this path to (ref: minivectordb/
vector_database.py:43-46, 475, 497, 511-514), is an un-vendored, un-pinned
dependency that cannot be installed in this image, and the reference's tests
hold no numeric golden vector for the scan.

Three layers, each usable on its own:

* ``IndexFlatIP`` / ``normalize_L2`` -- faiss-shaped objects backed by the C
  restatement (liboracle.so).  With ``install_as_faiss()`` the reference's own
  Python modules can be imported on top of them (used by
  tests/golden/make_golden.py to generate fixtures from the reference code).
* ``gold_scores`` / ``gold_topk`` -- float64 numpy scorer with a deterministic
  order (score descending, row ascending); the arbiter of near-ties.
* ``classify_parity`` -- the parity rule of SURVEY.md section 8c: ids equal
  position-wise, a mismatch excused only when the float64 scores of the two
  ids are closer than the fp32 accumulation bound; distances within 1e-5
  relative.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs import this.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import sys
import types

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

FLT_LOWEST = float(np.finfo(np.float32).min)


def build(force: bool = False) -> str:
    """Compile liboracle.so with oracle/Makefile (gcc is in the image)."""
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "faiss_flat_ip.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s", "-B", "liboracle.so"], check=True,
                       env={**os.environ, "CC": "/usr/bin/gcc"})
    return so


def _lib():
    global _LIB
    if _LIB is None:
        so = build()
        try:
            lib = ctypes.CDLL(so)
        except OSError:
            lib = ctypes.CDLL(build(force=True))
        f32p = ctypes.POINTER(ctypes.c_float)
        i64p = ctypes.POINTER(ctypes.c_int64)
        i64 = ctypes.c_int64
        lib.orc_renorm_L2.argtypes = [i64, i64, f32p]
        lib.orc_renorm_L2.restype = None
        lib.orc_search_flat_ip.argtypes = [f32p, i64, i64, f32p, i64, i64, f32p, i64p, ctypes.c_int]
        lib.orc_search_flat_ip.restype = ctypes.c_int
        lib.orc_search_gathered.argtypes = [f32p, i64, i64, i64p, i64, f32p, i64, i64, f32p, i64p,
                                            f32p, ctypes.c_int]
        lib.orc_search_gathered.restype = ctypes.c_int
        lib.orc_synth_rows.argtypes = [ctypes.c_uint64, i64, i64, i64, ctypes.c_int, f32p]
        lib.orc_synth_rows.restype = None
        lib.orc_num_threads.restype = ctypes.c_int
        lib.orc_reservoir_capacity.argtypes = [i64]
        lib.orc_reservoir_capacity.restype = i64
        lib.orc_block_add.argtypes = [ctypes.c_int, i64, i64, i64, i64, i64, f32p, f32p, i64p, i64p, f32p]
        lib.orc_block_add.restype = ctypes.c_int
        lib.orc_block_end.argtypes = [ctypes.c_int, i64, i64, i64, f32p, i64p, i64p, f32p, f32p, i64p]
        lib.orc_block_end.restype = ctypes.c_int
        _LIB = lib
    return _LIB


def _f32p(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def _i64p(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int64))


def num_threads() -> int:
    return int(_lib().orc_num_threads())


# ---------------------------------------------------------------------------
# faiss-shaped layer
# ---------------------------------------------------------------------------

def normalize_L2(x: np.ndarray) -> None:
    """faiss.normalize_L2: in place, float32 C-contiguous [n, d]."""
    if not (isinstance(x, np.ndarray) and x.dtype == np.float32 and x.ndim == 2
            and x.flags.c_contiguous):
        raise TypeError("normalize_L2 needs a C-contiguous float32 [n, d] array")
    n, d = x.shape
    _lib().orc_renorm_L2(d, n, _f32p(x))


def search_flat_ip(x: np.ndarray, q: np.ndarray, k: int, nthreads: int = 0):
    """IndexFlatIP.search on an explicit matrix: returns (D[nq,k], I[nq,k])."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    q = np.ascontiguousarray(q, dtype=np.float32)
    assert x.ndim == 2 and q.ndim == 2 and x.shape[1] == q.shape[1]
    assert k > 0
    nq = q.shape[0]
    D = np.empty((nq, k), dtype=np.float32)
    I = np.empty((nq, k), dtype=np.int64)
    rc = _lib().orc_search_flat_ip(_f32p(x), x.shape[0], x.shape[1], _f32p(q), nq, k,
                                   _f32p(D), _i64p(I), nthreads)
    if rc != 0:
        raise RuntimeError(f"orc_search_flat_ip rc={rc}")
    return D, I


BLAS_THRESHOLD = 20          # faiss distance_compute_blas_threshold: nq >= 20 takes the sgemm path
BLAS_QUERY_BS, BLAS_DB_BS = 4096, 1024   # distance_compute_blas_query_bs / _database_bs


class BlockHandler:
    """faiss's Top1 / Heap / Reservoir BlockResultHandler (k == 1 / k < 100 / k >= 100) for `nq` queries,
    fed block by block with inner products -- the consumer of the sgemm path."""

    def __init__(self, nq: int, k: int):
        self.nq, self.k = int(nq), int(k)
        self.kind = 0 if k == 1 else (1 if k < 100 else 2)
        self.cap = 1 if k == 1 else (k if k < 100 else int(_lib().orc_reservoir_capacity(k)))
        self.vals = np.full((nq, self.cap), FLT_LOWEST, dtype=np.float32)
        self.ids = np.full((nq, self.cap), -1, dtype=np.int64)
        self.cnt = np.zeros(nq, dtype=np.int64)
        self.thr = np.full(nq, FLT_LOWEST, dtype=np.float32)

    def add_results(self, j0: int, j1: int, ip: np.ndarray) -> None:
        ip = np.ascontiguousarray(ip, dtype=np.float32)
        assert ip.shape == (self.nq, j1 - j0)
        rc = _lib().orc_block_add(self.kind, self.k, self.cap, self.nq, j0, j1, _f32p(ip), _f32p(self.vals),
                                  _i64p(self.ids), _i64p(self.cnt), _f32p(self.thr))
        assert rc == 0

    def end(self):
        D = np.empty((self.nq, self.k), dtype=np.float32)
        I = np.empty((self.nq, self.k), dtype=np.int64)
        rc = _lib().orc_block_end(self.kind, self.k, self.cap, self.nq, _f32p(self.vals), _i64p(self.ids),
                                  _i64p(self.cnt), _f32p(self.thr), _f32p(D), _i64p(I))
        assert rc == 0
        return D, I


def search_flat_ip_blas(x: np.ndarray, q: np.ndarray, k: int, make_chunk=None, n: int = 0):
    """IndexFlatIP.search the way faiss runs it for nq >= 20 (exhaustive_inner_product_blas): for every
    block of 4096 queries, for every block of 1024 rows, one sgemm (here numpy's float32 matmul = the host
    BLAS, as faiss calls sgemm_) and the block of inner products goes to the block result handler.
    `x` may be None with make_chunk(row0, m) + n: the matrix is then produced 1M rows at a time."""
    q = np.ascontiguousarray(q, dtype=np.float32)
    nq = q.shape[0]
    n = x.shape[0] if x is not None else int(n)
    D = np.empty((nq, k), dtype=np.float32)
    I = np.empty((nq, k), dtype=np.int64)
    for i0 in range(0, nq, BLAS_QUERY_BS):
        i1 = min(nq, i0 + BLAS_QUERY_BS)
        h = BlockHandler(i1 - i0, k)
        step = 1 << 20 if x is None else n
        for c0 in range(0, n, max(step, 1)):
            c1 = min(n, c0 + step)
            xc = x if x is not None else make_chunk(c0, c1 - c0)
            for j0 in range(c0, c1, BLAS_DB_BS):
                j1 = min(c1, j0 + BLAS_DB_BS)
                h.add_results(j0, j1, q[i0:i1] @ xc[j0 - c0:j1 - c0].T)
        D[i0:i1], I[i0:i1] = h.end()
    return D, I


def search_gathered(x: np.ndarray, rows: np.ndarray, q: np.ndarray, k: int, nthreads: int = 0):
    """The reference's filtered branch (VDB:508-523): gather `rows` (in the
    order given) into a temporary index and search it.  Returns (D, P) with P
    positions into `rows`."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    q = np.ascontiguousarray(q, dtype=np.float32)
    rows = np.ascontiguousarray(rows, dtype=np.int64)
    nq = q.shape[0]
    D = np.empty((nq, k), dtype=np.float32)
    P = np.empty((nq, k), dtype=np.int64)
    scratch = np.empty((max(len(rows), 1), x.shape[1]), dtype=np.float32)
    rc = _lib().orc_search_gathered(_f32p(x), x.shape[0], x.shape[1], _i64p(rows), len(rows),
                                    _f32p(q), nq, k, _f32p(D), _i64p(P), _f32p(scratch), nthreads)
    if rc != 0:
        raise RuntimeError(f"orc_search_gathered rc={rc}")
    return D, P


def search_masked(x: np.ndarray, admissible: np.ndarray, q: np.ndarray, k: int, nthreads: int = 0):
    """Bit-mask flavour of the filtered branch: admissible rows gathered in
    ascending row order, results mapped back to row numbers (-1 padded)."""
    rows = np.flatnonzero(np.asarray(admissible, dtype=bool)).astype(np.int64)
    D, P = search_gathered(x, rows, q, k, nthreads)
    I = np.where(P >= 0, rows[np.clip(P, 0, max(len(rows) - 1, 0))] if len(rows) else -1, -1)
    return D, I.astype(np.int64)


class IndexFlatIP:
    """faiss.IndexFlatIP restated: keeps a private copy of the rows it is given
    (faiss copies on add; ref VDB:46) and scans them on search (VDB:497)."""

    def __init__(self, d: int):
        self.d = int(d)
        self._chunks = []
        self._x = np.zeros((0, self.d), dtype=np.float32)
        self.ntotal = 0

    def add(self, x):
        x = np.ascontiguousarray(x, dtype=np.float32)
        assert x.ndim == 2 and x.shape[1] == self.d
        self._chunks.append(x.copy())
        self.ntotal += x.shape[0]

    def _matrix(self):
        if self._chunks:
            self._x = np.vstack([self._x] + self._chunks)
            self._chunks = []
        return self._x

    def search(self, q, k, nthreads: int = 0):
        q = np.ascontiguousarray(q, dtype=np.float32)
        assert q.ndim == 2 and q.shape[1] == self.d
        assert k > 0
        return search_flat_ip(self._matrix(), q, int(k), nthreads)


def install_as_faiss() -> types.ModuleType:
    """Register a module named ``faiss`` backed by this oracle so that the
    reference's Python files (which do ``import faiss``) run unmodified.
    Used only by tests/golden/make_golden.py inside the build container."""
    m = types.ModuleType("faiss")
    m.IndexFlatIP = IndexFlatIP
    m.normalize_L2 = normalize_L2
    m.__oracle__ = True
    sys.modules["faiss"] = m
    return m


# ---------------------------------------------------------------------------
# synthetic data (bit-identical to the CUDA generator in csrc/)
# ---------------------------------------------------------------------------

DIST_BELL, DIST_UNIFORM = 0, 1


def synth_rows(seed: int, row0: int, n: int, d: int, dist: int = DIST_BELL, out=None) -> np.ndarray:
    """`out`: optional float32 C-contiguous buffer of at least n*d elements to generate into (a
    streamed scan reuses one chunk buffer instead of faulting in fresh pages per chunk)."""
    if out is None:
        out = np.empty((n, d), dtype=np.float32)
    else:
        assert out.dtype == np.float32 and out.flags.c_contiguous and out.size >= n * d
        out = out.reshape(-1)[:n * d].reshape(n, d)
    _lib().orc_synth_rows(seed, row0, n, d, dist, _f32p(out))
    return out


def synth_rows_numpy(seed: int, row0: int, n: int, d: int, dist: int = DIST_BELL) -> np.ndarray:
    """Same generator in pure numpy (cross-check of the C and CUDA versions)."""
    M = np.uint64
    with np.errstate(over="ignore"):
        rows = (np.arange(row0, row0 + n, dtype=np.uint64)[:, None] << M(20))
        z = rows + np.arange(d, dtype=np.uint64)[None, :] + M(seed) * M(0x9E3779B97F4A7C15)
        z = (z ^ (z >> M(30))) * M(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> M(27))) * M(0x94D049BB133111EB)
        z = z ^ (z >> M(31))
    if dist == DIST_BELL:
        s = ((z & M(0xFFFF)) + ((z >> M(16)) & M(0xFFFF)) + ((z >> M(32)) & M(0xFFFF))
             + (z >> M(48))).astype(np.int64)
        return ((s - 131070).astype(np.float32) * np.float32(1.0 / 65536.0)).astype(np.float32)
    return ((z >> M(40)).astype(np.float32) * np.float32(1.0 / 16777216.0)).astype(np.float32)


# ---------------------------------------------------------------------------
# float64 gold + parity rule
# ---------------------------------------------------------------------------

def gold_scores(x: np.ndarray, q: np.ndarray) -> np.ndarray:
    """float64 inner products [nq, n] of the float32 inputs."""
    return np.asarray(q, dtype=np.float64) @ np.asarray(x, dtype=np.float64).T


def gold_topk(x: np.ndarray, q: np.ndarray, k: int, admissible=None):
    """Deterministic float64 top-k: score descending, row ascending; -1 padded."""
    S = gold_scores(x, q)
    nq, n = S.shape
    if admissible is not None:
        S = np.where(np.asarray(admissible, dtype=bool)[None, :], S, -np.inf)
    D = np.full((nq, k), FLT_LOWEST, dtype=np.float64)
    I = np.full((nq, k), -1, dtype=np.int64)
    for i in range(nq):
        order = np.lexsort((np.arange(n), -S[i]))[:k]
        order = order[np.isfinite(S[i][order])]
        D[i, :len(order)] = S[i][order]
        I[i, :len(order)] = order
    return D, I


def classify_parity(x, q, I_test, D_test, I_ref, D_ref, rel_tol=1e-5, admissible=None):
    """Apply the parity rule.  Returns a dict with counts of position-wise id
    matches, mismatches excused as exact ties / fp near-ties (judged in
    float64), real errors, and the worst relative distance error.

    The fp32 accumulation bound used for a near-tie is
    |s_a - s_b| <= 4 * d * 2^-24 * ||q|| * max||x||  (loose, order-independent)."""
    x = np.asarray(x, dtype=np.float32)
    q = np.asarray(q, dtype=np.float32)
    nq, k = I_ref.shape
    d = x.shape[1]
    xn = float(np.sqrt((x.astype(np.float64) ** 2).sum(1)).max()) if len(x) else 0.0
    out = dict(positions=int(nq * k), id_equal=0, exact_tie=0, near_tie=0, real_error=0,
               max_rel_err=0.0, set_equal_queries=0)
    for i in range(nq):
        qi = q[i].astype(np.float64)
        bound = 4.0 * d * 2.0 ** -24 * float(np.linalg.norm(qi)) * xn
        if set(I_test[i].tolist()) == set(I_ref[i].tolist()):
            out["set_equal_queries"] += 1
        for j in range(k):
            a, b = int(I_test[i, j]), int(I_ref[i, j])
            if a == b:
                out["id_equal"] += 1
            elif a < 0 or b < 0:
                out["real_error"] += 1
                continue
            else:
                sa = float(x[a].astype(np.float64) @ qi)
                sb = float(x[b].astype(np.float64) @ qi)
                if admissible is not None and not (admissible[a] and admissible[b]):
                    out["real_error"] += 1
                elif sa == sb:
                    out["exact_tie"] += 1
                elif abs(sa - sb) <= bound:
                    out["near_tie"] += 1
                else:
                    out["real_error"] += 1
            if b >= 0 and a >= 0:
                ref = float(D_ref[i, j])
                # relative error; scores are cosines, so |ref| is floored at 1 % of
                # ||q||*max||x|| to keep the rule meaningful for near-zero scores
                floor = 0.01 * float(np.linalg.norm(qi)) * xn
                err = abs(float(D_test[i, j]) - ref) / max(abs(ref), floor, 1e-30)
                out["max_rel_err"] = max(out["max_rel_err"], err)
    out["ok"] = out["real_error"] == 0 and out["max_rel_err"] <= rel_tol
    return out


# ---------------------------------------------------------------------------
# full-size checks: the matrix does not fit (or is not worth keeping) in host memory
# ---------------------------------------------------------------------------

def merge_topk_lists(parts, k):
    """Merge per-chunk / per-shard (D [nq,kk], I [nq,kk]) lists whose labels are already global:
    score descending, label ascending on exact ties (what ONE sequential scan over all rows with a
    strict '>' keeps); -1 / -FLT_MAX padded."""
    nq = parts[0][0].shape[0]
    D = np.full((nq, k), FLT_LOWEST, dtype=np.float32)
    I = np.full((nq, k), -1, dtype=np.int64)
    for i in range(nq):
        d = np.concatenate([p[0][i] for p in parts])
        l = np.concatenate([p[1][i] for p in parts])
        keep = l >= 0
        d, l = d[keep], l[keep]
        order = np.lexsort((l, -d.astype(np.float64)))[:k]
        D[i, :len(order)] = d[order]
        I[i, :len(order)] = l[order]
    return D, I


def search_streamed(make_chunk, n, q, k, chunk_rows=1 << 20, row_offset=0, admissible=None, nthreads=0):
    """IndexFlatIP.search over a matrix that is PRODUCED chunk by chunk: make_chunk(row0, m) returns
    rows [row0, row0+m) as float32 [m, d] (already normalised if the index holds normalised rows).
    Every chunk is scanned by the C restatement and the per-chunk top-k are merged on the host.
    Labels are row numbers + row_offset.  `admissible` (bool[n]) = the filtered branch."""
    parts = []
    for r0 in range(0, n, chunk_rows):
        m = min(chunk_rows, n - r0)
        x = make_chunk(r0, m)
        if admissible is None:
            D, I = search_flat_ip(x, q, k, nthreads)
        else:
            D, I = search_masked(x, admissible[r0:r0 + m], q, k, nthreads)
        parts.append((D, np.where(I >= 0, I + r0 + row_offset, -1)))
        if len(parts) >= 8:
            parts = [merge_topk_lists(parts, k)]
    if not parts:
        nq = q.shape[0]
        return np.full((nq, k), FLT_LOWEST, dtype=np.float32), np.full((nq, k), -1, dtype=np.int64)
    return merge_topk_lists(parts, k)


def classify_parity_lazy(fetch_row, d, q, I_test, D_test, I_ref, D_ref, rel_tol=1e-5, max_norm=1.0,
                         admissible=None):
    """`classify_parity` for matrices that are not in memory: fetch_row(label) -> float32 [d] is only
    called for the two ids of a position-wise mismatch.  `max_norm` = largest row norm (1 for a
    normalised index).  `admissible(label) -> bool` optional."""
    q = np.asarray(q, dtype=np.float32)
    nq, k = I_ref.shape
    out = dict(positions=int(nq * k), id_equal=0, exact_tie=0, near_tie=0, real_error=0,
               max_rel_err=0.0, set_equal_queries=0)
    for i in range(nq):
        qi = q[i].astype(np.float64)
        qn = float(np.linalg.norm(qi))
        bound = 4.0 * d * 2.0 ** -24 * qn * max_norm
        if set(I_test[i].tolist()) == set(I_ref[i].tolist()):
            out["set_equal_queries"] += 1
        for j in range(k):
            a, b = int(I_test[i, j]), int(I_ref[i, j])
            if a == b:
                out["id_equal"] += 1
            elif a < 0 or b < 0:
                out["real_error"] += 1
                continue
            else:
                sa = float(np.asarray(fetch_row(a), dtype=np.float64) @ qi)
                sb = float(np.asarray(fetch_row(b), dtype=np.float64) @ qi)
                if admissible is not None and not (admissible(a) and admissible(b)):
                    out["real_error"] += 1
                elif sa == sb:
                    out["exact_tie"] += 1
                elif abs(sa - sb) <= bound:
                    out["near_tie"] += 1
                else:
                    out["real_error"] += 1
            if a >= 0 and b >= 0:
                ref = float(D_ref[i, j])
                err = abs(float(D_test[i, j]) - ref) / max(abs(ref), 0.01 * qn * max_norm, 1e-30)
                out["max_rel_err"] = max(out["max_rel_err"], err)
    out["ok"] = out["real_error"] == 0 and out["max_rel_err"] <= rel_tol
    return out
