"""GPU: parity with the CPU oracle at BASELINE.json's FULL sizes.  The oracle (C restatement of faiss's
flat-IP search) is streamed over 1M-row chunks of the same counter-generated matrix, so nothing has to fit
in host memory; the parity rule is oracle.classify_parity's (ids position-wise, a mismatch excused only for
exact / fp32-near ties judged in float64, distances <= 1e-5 relative).  Every test prints its tie counts."""
import threading

import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu
SEED_DB, SEED_Q = 1234, 4321


@pytest.fixture(scope="module")
def mv():
    import minivectordb_b200 as m
    return m


def _need_hbm(gb):
    import torch
    free_b, _ = torch.cuda.mem_get_info()
    if free_b < gb * (1 << 30):
        pytest.skip(f"needs ~{gb} GB of free HBM")


def _chunker(d, seed=SEED_DB, row_of=None):
    buf = np.empty((1 << 20, d), dtype=np.float32)

    def chunk(r0, m):
        x = O.synth_rows(seed, r0, m, d, out=buf)
        O.normalize_L2(x)
        return x
    return chunk


def _fetch(d, seed=SEED_DB):
    def fetch(label):
        x = O.synth_rows(seed, int(label), 1, d)
        O.normalize_L2(x)
        return x[0]
    return fetch


def _report(name, rep):
    print(f"[fullsize] {name}: ids_equal={rep['id_equal']}/{rep['positions']} exact_ties={rep['exact_tie']} "
          f"near_ties={rep['near_tie']} real_errors={rep['real_error']} max_rel_err={rep['max_rel_err']:.2e}")
    assert rep["ok"], (name, rep)


def test_config2_all_rows_against_the_oracle(mv):
    """C2: 1M x 384, k = 10 (and 100), ~50 % filter -- the oracle scans ALL 1M rows."""
    n, d = 1_000_000, 384
    eng = mv.FlatIPEngine(d)
    eng.add_synthetic(SEED_DB, 0, n, dist=0, normalize=True)
    q = O.synth_rows(SEED_Q, 0, 8, d)
    O.normalize_L2(q)
    adm = O.synth_rows(99, 0, 1, n, O.DIST_UNIFORM)[0] > 0.5
    x = O.synth_rows(SEED_DB, 0, n, d)
    O.normalize_L2(x)
    for k in (10, 100):
        D, I = eng.search(q, k, mask=adm)
        Dr, Ir = O.search_masked(x, adm, q, k)
        _report(f"C2 filtered k={k}", O.classify_parity(x, q, I, D, Ir, Dr, admissible=adm))
        D, I = eng.search(q, k)
        Dr, Ir = O.search_flat_ip(x, q, k)
        _report(f"C2 unfiltered k={k}", O.classify_parity(x, q, I, D, Ir, Dr))
    for i in range(4):   # one query at a time: the single-query kernel
        D, I = eng.search(q[i:i + 1], 10, mask=adm)
        Dr, Ir = O.search_masked(x, adm, q[i:i + 1], 10)
        assert O.classify_parity(x, q[i:i + 1], I, D, Ir, Dr, admissible=adm)["ok"]
    eng.close()


def test_config4_shard_against_the_streamed_oracle(mv):
    """C4 shard: 12.5M x 512 (25.6 GB), single queries, k = 10, unfiltered."""
    _need_hbm(40)
    n, d, k = 12_500_000, 512, 10
    eng = mv.FlatIPEngine(d, capacity_hint=n)
    eng.add_synthetic(SEED_DB, 0, n, dist=0, normalize=True)
    q = O.synth_rows(SEED_Q, 0, 4, d)
    O.normalize_L2(q)
    D = np.empty((4, k), dtype=np.float32)
    I = np.empty((4, k), dtype=np.int64)
    for i in range(4):
        D[i:i + 1], I[i:i + 1] = eng.search(q[i:i + 1], k)
    Dr, Ir = O.search_streamed(_chunker(d), n, q, k, nthreads=4)
    _report("C4 shard", O.classify_parity_lazy(_fetch(d), d, q, I, D, Ir, Dr))
    eng.close()


def test_config3_against_the_blocked_sgemm_oracle(mv):
    """C3: 10M x 1024, 4096 queries, k = 100 on the tensor-core path (exact mode), checked on 64 of the
    queries against the oracle's nq >= 20 path: 1024-row sgemm blocks feeding faiss's ReservoirTopN
    (capacity (2k+15)&~15, partition_fuzzy)."""
    _need_hbm(80)
    n, d, nq, k = 10_000_000, 1024, 4096, 100
    eng = mv.FlatIPEngine(d, capacity_hint=n)
    eng.add_synthetic(SEED_DB, 0, n, dist=0, normalize=True)
    q = O.synth_rows(SEED_Q, 0, nq, d)
    O.normalize_L2(q)
    D, I = eng.search(q, k)
    sample = np.arange(0, nq, 64)          # 64 queries
    Dr, Ir = O.search_flat_ip_blas(None, q[sample], k, make_chunk=_chunker(d), n=n)
    _report("C3 exact vs blocked-sgemm oracle", O.classify_parity_lazy(_fetch(d), d, q[sample], I[sample], D[sample], Ir, Dr))
    eng.set_option("batch_mode", 2)        # bf16 mode: recall against the oracle
    Db, Ib = eng.search(q[sample], k)
    recall = float(np.mean([len(set(Ib[i]) & set(Ir[i])) / k for i in range(len(sample))]))
    print(f"[fullsize] C3 bf16 recall@100 vs oracle = {recall:.4f}")
    assert recall > 0.98
    eng.close()


def test_config5_after_churn_against_the_streamed_oracle(mv):
    """C5 shape: 10M x 768 with deletes, inserts and concurrent filtered / unfiltered query threads;
    then, QUIESCED (the reference is not linearizable under churn either, ref vector_database.py:497-507),
    the tombstone + mask path is compared with the oracle over the surviving rows -- before and after the
    order-preserving compaction."""
    _need_hbm(60)
    n0, extra, d, k = 10_000_000, 200_000, 768, 10
    eng = mv.FlatIPEngine(d, capacity_hint=n0 + extra)
    eng.add_synthetic(SEED_DB, 0, n0, dist=0, normalize=True)
    rng = np.random.default_rng(7)
    q = O.synth_rows(SEED_Q, 0, 8, d)
    O.normalize_L2(q)
    dead = rng.choice(n0, size=500_000, replace=False)
    filt = O.synth_rows(99, 0, 1, n0 + extra, O.DIST_UNIFORM)[0] > 0.5
    stop = threading.Event()
    errs = []

    def searcher(i):
        try:
            while not stop.is_set():
                D, I = eng.search(q[i % 8:i % 8 + 1], k, mask=filt[:eng.ntotal] if i % 2 else None)
                assert I[0, 0] >= 0 and np.all(np.diff(D[0]) <= 0)
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    ts = [threading.Thread(target=searcher, args=(i,)) for i in range(4)]
    [t.start() for t in ts]
    for part in np.array_split(dead, 20):                      # deletes ...
        eng.remove_rows(part)
    for j in range(10):                                        # ... and inserts while queries run
        eng.add_synthetic(SEED_DB, n0 + j * (extra // 10), extra // 10, dist=0, normalize=True)
    stop.set()
    [t.join() for t in ts]
    assert not errs, errs[:1]
    n = n0 + extra
    assert eng.ntotal == n and eng.nlive == n - dead.size
    live = np.ones(n, dtype=bool)
    live[dead] = False
    for name, adm in (("tombstones", live), ("tombstones + filter", live & filt)):
        D, I = eng.search(q[:4], k, mask=None if adm is live else filt)
        Dr, Ir = O.search_streamed(_chunker(d), n, q[:4], k, admissible=adm, nthreads=4)
        _report(f"C5 quiesced, {name}", O.classify_parity_lazy(_fetch(d), d, q[:4], I, D, Ir, Dr, admissible=lambda r: bool(adm[r])))
    # compaction squeezes the tombstones out, order preserved: row r becomes rank(r) among the live rows
    assert eng.compact() == n - dead.size
    rank = np.cumsum(live) - 1
    D, I = eng.search(q[:4], k)
    Dr, Ir = O.search_streamed(_chunker(d), n, q[:4], k, admissible=live, nthreads=4)
    Ir_c = np.where(Ir >= 0, rank[np.clip(Ir, 0, n - 1)], -1)
    back = np.flatnonzero(live)
    _report("C5 after compaction", O.classify_parity_lazy(lambda r: _fetch(d)(back[r]), d, q[:4], I, D, Ir_c, Dr))
    eng.close()
