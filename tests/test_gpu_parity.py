"""GPU parity: the CUDA path (through the C ABI) against the oracle on the same
seeded inputs.  Bar (BASELINE.json north_star): ids equal position-wise except
exact / fp near-ties (judged in float64), distances within 1e-5 relative."""
import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu

REL_TOL = 1e-5  # north_star: "distances must agree within 1e-5 relative"


@pytest.fixture(scope="module")
def mv():
    import minivectordb_b200 as m
    return m


def _data(n, d, nq, seed=1, dist=O.DIST_BELL):
    x = O.synth_rows(seed, 0, n, d, dist)
    O.normalize_L2(x)
    q = O.synth_rows(seed + 100, 0, nq, d, dist)
    O.normalize_L2(q)
    return x, q


def _check(x, q, k, D, I, adm=None):
    if adm is None:
        Dr, Ir = O.search_flat_ip(x, q, k)
    else:
        Dr, Ir = O.search_masked(x, adm, q, k)
    rep = O.classify_parity(x, q, I, D, Ir, Dr, rel_tol=REL_TOL, admissible=adm)
    assert rep["ok"], rep
    # padding identical to faiss
    assert np.array_equal(I < 0, Ir < 0)
    assert np.all(D[I < 0] == np.finfo(np.float32).min)
    return rep


@pytest.mark.parametrize("variant", [1, 2])  # TMA ring, direct LDG
@pytest.mark.parametrize("n,d", [(1, 2), (7, 2), (100, 3), (1000, 32), (5000, 64), (4097, 100), (20000, 384),
                                 (12345, 512), (3000, 768), (2000, 1024), (600, 1536), (300, 4096)])
def test_single_query_topk(mv, variant, n, d):
    x, q = _data(n, d, 3, seed=n + d)
    eng = mv.FlatIPEngine(d)
    eng.set_option("scan_variant", variant)
    eng.add(x)
    for k in (1, 10, 100):
        D, I = eng.search(q, k)
        _check(x, q, k, D, I)
    eng.close()


@pytest.mark.parametrize("variant", [1, 2])
@pytest.mark.parametrize("nq", [2, 3, 4, 5, 8, 11, 19])
def test_small_batches(mv, variant, nq):
    x, q = _data(9000, 384, nq, seed=nq)
    eng = mv.FlatIPEngine(384)
    eng.set_option("scan_variant", variant)
    eng.add(x)
    D, I = eng.search(q, 10)
    _check(x, q, 10, D, I)
    # a batch must give exactly what the same queries give one by one
    for i in range(nq):
        D1, I1 = eng.search(q[i:i + 1], 10)
        assert np.array_equal(I1[0], I[i]) and np.array_equal(D1[0], D[i])
    eng.close()


@pytest.mark.parametrize("variant", [1, 2])
@pytest.mark.parametrize("frac", [0.0, 0.001, 0.12, 0.5, 1.0])
def test_bitmask_filter(mv, variant, frac):
    n, d = 30011, 384
    x, q = _data(n, d, 4, seed=5)
    adm = np.random.default_rng(int(frac * 1000)).random(n) < frac
    eng = mv.FlatIPEngine(d)
    eng.set_option("scan_variant", variant)
    eng.add(x)
    for k in (10, 100):
        D, I = eng.search(q, k, mask=adm)
        _check(x, q, k, D, I, adm)
    eng.close()


def test_mask_shorter_than_index_hides_new_rows(mv):
    x, q = _data(1000, 64, 2)
    eng = mv.FlatIPEngine(64)
    eng.add(x[:600])
    adm = np.ones(600, dtype=bool)
    eng.add(x[600:])  # appended after the caller built its mask
    D, I = eng.search(q, 10, mask=adm)
    _check(x[:600], q, 10, D, I)
    eng.close()


@pytest.mark.parametrize("k", [129, 500, 825, 999, 5000])
def test_large_k_path(mv, k):
    # reference tests search with k = 500 / 825 / 999
    n, d = 6000, 64
    x, q = _data(n, d, 2, seed=k)
    eng = mv.FlatIPEngine(d)
    eng.add(x)
    D, I = eng.search(q, k)
    _check(x, q, k, D, I)
    adm = np.random.default_rng(1).random(n) < 0.1  # fewer admissible rows than k -> padding
    D, I = eng.search(q, k, mask=adm)
    _check(x, q, k, D, I, adm)
    eng.close()


def test_k_larger_than_n_and_empty_index(mv):
    x, q = _data(5, 16, 2)
    eng = mv.FlatIPEngine(16)
    D, I = eng.search(q, 4)  # empty index
    assert np.all(I == -1) and np.all(D == np.finfo(np.float32).min)
    eng.add(x)
    D, I = eng.search(q, 8)
    _check(x, q, 8, D, I)
    D, I = eng.search(q, 300)  # large-k path with k > n
    _check(x, q, 300, D, I)
    eng.close()


def test_duplicate_vectors_tie_order_is_row_ascending(mv):
    x, q = _data(64, 32, 1)
    xd = np.repeat(x[:4], 50, axis=0)  # rows 0..49 identical, 50..99 identical, ...
    eng = mv.FlatIPEngine(32)
    eng.add(xd)
    D, I = eng.search(q, 20)
    _check(xd, q, 20, D, I)
    # inside a run of equal scores rows ascend (the engine's documented tie rule)
    for a in range(19):
        if D[0, a] == D[0, a + 1]:
            assert I[0, a] < I[0, a + 1]
    eng.close()


def test_uniform_positive_vectors_like_the_reference_tests(mv):
    # np.random.rand vectors: all scores crowd near 0.75 (ref tests/test_multithreaded_operations.py:13)
    x, q = _data(20000, 64, 4, seed=9, dist=O.DIST_UNIFORM)
    eng = mv.FlatIPEngine(64)
    eng.add(x)
    D, I = eng.search(q, 10)
    _check(x, q, 10, D, I)
    eng.close()


def test_engine_side_normalisation_matches_oracle(mv):
    raw = O.synth_rows(21, 0, 4000, 384)
    raw[17] = 0.0
    qraw = O.synth_rows(22, 0, 3, 384) * 3.0
    eng = mv.FlatIPEngine(384)
    eng.add(raw, normalize=True)
    got = eng.reconstruct_n(0, 4000)
    ref = raw.copy()
    O.normalize_L2(ref)
    assert np.allclose(got, ref, rtol=0, atol=2e-7)
    assert np.all(got[17] == 0.0)
    # faiss.normalize_L2 drop-in on host data
    h = raw.copy()
    mv.normalize_L2(h)
    assert np.array_equal(h, got)
    # query normalisation fused into the scan (VDB:475)
    D, I = eng.search(qraw, 10, normalize=True)
    qn = qraw.copy()
    O.normalize_L2(qn)
    _check(got, qn, 10, D, I)
    eng.close()


def test_synthetic_generator_is_bit_identical_on_device(mv):
    eng = mv.FlatIPEngine(100)
    eng.add_synthetic(1234, 50, 3000, dist=0, normalize=False)
    assert np.array_equal(eng.reconstruct_n(0, 3000), O.synth_rows(1234, 50, 3000, 100))
    eng.add_synthetic(7, 0, 100, dist=1, normalize=False)
    assert np.array_equal(eng.reconstruct_n(3000, 100), O.synth_rows(7, 0, 100, 100, O.DIST_UNIFORM))
    eng.close()


def test_tombstones_and_order_preserving_compaction(mv):
    n, d = 10000, 64
    x, q = _data(n, d, 3, seed=3)
    eng = mv.FlatIPEngine(d)
    eng.add(x[:4000])
    eng.add(x[4000:])  # growth across two adds
    rng = np.random.default_rng(0)
    dead = rng.choice(n, 3000, replace=False)
    eng.remove_rows(dead)
    live = np.ones(n, dtype=bool)
    live[dead] = False
    assert (eng.ntotal, eng.nlive) == (n, n - 3000)
    D, I = eng.search(q, 10)
    _check(x, q, 10, D, I, live)
    # filter AND tombstones
    adm = rng.random(n) < 0.5
    D, I = eng.search(q, 10, mask=adm)
    _check(x, q, 10, D, I, adm & live)
    # errors: unknown / double delete leave the index untouched
    from minivectordb_b200._native import MvdbError
    with pytest.raises(MvdbError):
        eng.remove_rows([int(dead[0])])
    with pytest.raises(MvdbError):
        eng.remove_rows([n + 5])
    assert eng.nlive == n - 3000
    # compaction renumbers densely, order preserved (ref VDB:138-152)
    assert eng.compact() == n - 3000
    xs = x[live]
    assert np.array_equal(eng.reconstruct_n(0, n - 3000), xs)
    D, I = eng.search(q, 10)
    _check(xs, q, 10, D, I)
    eng.add(x[:10])
    assert eng.ntotal == n - 3000 + 10
    eng.close()


def test_reset_and_regrow(mv):
    x, q = _data(3000, 32, 2)
    eng = mv.FlatIPEngine(32)
    eng.add(x)
    eng.reset()
    assert eng.ntotal == 0
    eng.add(x[:100])
    D, I = eng.search(q, 5)
    _check(x[:100], q, 5, D, I)
    eng.close()


def test_argument_errors_raise_not_abort(mv):
    from minivectordb_b200._native import MvdbError
    eng = mv.FlatIPEngine(8)
    with pytest.raises(ValueError):
        eng.add(np.zeros((2, 9), dtype=np.float32))
    with pytest.raises(ValueError):
        eng.search(np.zeros((1, 8), dtype=np.float32), 0)
    with pytest.raises(MvdbError):
        eng.reconstruct(0)
    with pytest.raises(MvdbError):
        mv.FlatIPEngine(0)
    eng.close()


def test_faiss_shim_round_trip(mv):
    x, q = _data(2000, 48, 2)
    idx = mv.faiss_shim.IndexFlatIP(48)
    idx.add(x)
    assert idx.ntotal == 2000
    D, I = idx.search(q, 10)
    _check(x, q, 10, D, I)


def test_concurrent_searches_from_threads(mv):
    import threading
    x, q = _data(20000, 128, 16, seed=4)
    eng = mv.FlatIPEngine(128)
    eng.add(x)
    Dr, Ir = O.search_flat_ip(x, q, 10)
    errs = []

    def worker(i):
        try:
            for _ in range(50):
                D, I = eng.search(q[i:i + 1], 10)
                rep = O.classify_parity(x, q[i:i + 1], I, D, Ir[i:i + 1], Dr[i:i + 1])
                assert rep["ok"], rep
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    ts = [threading.Thread(target=worker, args=(i,)) for i in range(16)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errs, errs[:1]
    eng.close()


def test_merge_topk_device(mv):
    import torch
    x, q = _data(8000, 64, 5, seed=8)
    k, parts = 10, 4
    bounds = np.linspace(0, 8000, parts + 1).astype(int)
    Ds, Is = [], []
    for p in range(parts):
        eng = mv.FlatIPEngine(64)
        eng.add(x[bounds[p]:bounds[p + 1]])
        D, I = eng.search(q, k)
        Ds.append(D)
        Is.append(np.where(I >= 0, I + bounds[p], -1))
        eng.close()
    Dp = torch.tensor(np.stack(Ds)).cuda()
    Ip = torch.tensor(np.stack(Is)).cuda()
    Do = torch.empty((5, k), dtype=torch.float32, device="cuda")
    Io = torch.empty((5, k), dtype=torch.int64, device="cuda")
    mv.merge_topk_device(0, Dp.data_ptr(), Ip.data_ptr(), parts, 5, k, Do.data_ptr(), Io.data_ptr(),
                         torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    _check(x, q, k, Do.cpu().numpy(), Io.cpu().numpy())


def test_full_size_properties_config2(mv):
    """BASELINE config 2 at full size (1M x 384, k=10, ~50 % filter): the CPU
    oracle is checked on a slice; the whole run through size-independent
    properties: sortedness, admissibility, idempotence, and score re-derivation
    from reconstructed rows."""
    n, d, k = 1_000_000, 384, 10
    eng = mv.FlatIPEngine(d)
    eng.add_synthetic(1234, 0, n, dist=0, normalize=True)
    q = O.synth_rows(4321, 0, 4, d)
    O.normalize_L2(q)
    adm = O.synth_rows(99, 0, 1, n, O.DIST_UNIFORM)[0] > 0.5
    D, I = eng.search(q, k, mask=adm)
    assert np.all(np.diff(D, axis=1) <= 0)
    assert adm[I].all()
    D2, I2 = eng.search(q, k, mask=adm)
    assert np.array_equal(I, I2) and np.array_equal(D, D2)
    for qi in range(4):
        rows = np.stack([eng.reconstruct(int(r)) for r in I[qi]])
        s = rows.astype(np.float64) @ q[qi].astype(np.float64)
        assert np.allclose(s, D[qi], rtol=REL_TOL, atol=0)
    # oracle on the first 200k rows, engine restricted to them by the mask
    m = 200_000
    xs = eng.reconstruct_n(0, m)
    sub = adm.copy()
    sub[m:] = False
    D, I = eng.search(q, k, mask=sub)
    _check(xs, q, k, D, I, sub[:m])
    eng.close()


@pytest.mark.parametrize("d", [384, 512, 768, 1024, 1536])
def test_tma_ring_any_consumer_warp_count(mv, d):
    """Regression: the ring depth must be a multiple of the consumer-warp count
    (a warp running ahead must never pass a full-barrier wait on an older
    phase's parity).  Many tiles per CTA so that warps do drift apart."""
    n = 120_000
    eng = mv.FlatIPEngine(d)
    eng.add_synthetic(7, 0, n, dist=0, normalize=True)
    q = O.synth_rows(8, 0, 2, d)
    O.normalize_L2(q)
    eng.set_option("scan_variant", 2)
    Dref, Iref = eng.search(q, 10)
    eng.set_option("scan_variant", 1)
    for cw in range(1, 9):
        eng.set_option("consumer_warps", cw)
        for _ in range(3):
            D, I = eng.search(q, 10)
            assert np.array_equal(I, Iref) and np.array_equal(D, Dref), (d, cw)
    eng.close()


@pytest.mark.parametrize("n,d,grid", [(120_000, 384, 0), (99_991, 512, 0), (120_000, 200, 37), (5_000, 768, 0),
                                      (131, 64, 0), (64_123, 1536, 146)])
def test_dynamic_tile_schedule_is_result_neutral(mv, n, d, grid):
    """The TMA scan serves part of the tiles from a global counter ("dyn_tiles" = that percentage).
    Which CTA scans which tile must not show in the results: every percentage, with masks,
    tombstones, 1/2/4 queries per launch, ragged last tile/word and a grid that is not a multiple
    of 4, returns exactly what the static split (0) and the oracle return."""
    eng = mv.FlatIPEngine(d)
    eng.add_synthetic(21, 0, n, dist=0, normalize=True)
    eng.set_option("batch_mode", 0)
    eng.set_option("scan_variant", 1)
    if grid:
        eng.set_option("grid_ctas", grid)
    eng.remove_rows(np.arange(3, n, 11))
    adm = np.random.default_rng(22).random(n) < 0.5
    q = O.synth_rows(23, 0, 4, d)
    O.normalize_L2(q)
    x = O.synth_rows(21, 0, n, d)
    O.normalize_L2(x)
    live = np.ones(n, dtype=bool)
    live[np.arange(3, n, 11)] = False
    ref = {}
    for pct in (0, 15, 50, 100):
        eng.set_option("dyn_tiles", pct)
        for nq in (1, 2, 4):
            for mask in (None, adm):
                for rep in range(2):   # the counter must come back to zero after every launch
                    D, I = eng.search(q[:nq], 10, mask=mask)
                    key = (nq, mask is None)
                    if key not in ref:
                        ref[key] = (D, I)
                        _check(x, q[:nq], 10, D, I, live if mask is None else (live & adm))
                    assert np.array_equal(I, ref[key][1]) and np.array_equal(D, ref[key][0]), (pct, nq, mask is None, rep)
    eng.close()


@pytest.mark.parametrize("n,d", [(3_000, 64), (40_000, 384), (150_000, 512), (20_000, 1024)])
def test_programmatic_dependent_launch_is_result_neutral(mv, n, d):
    """Option "pdl": searches enqueued back to back on one stream overlap (the next scan starts while
    the previous one is still merging).  A shuffled mix of launches -- 1/2/4 queries, k = 10 and 100
    (different kernels, shared-memory footprints and trigger points), with a filter mask, three
    rounds without any synchronisation in between -- must return exactly what plain stream order
    returns and match the oracle; and when every search writes the SAME output slot the last one
    enqueued must be the one that stays."""
    import torch
    eng = mv.FlatIPEngine(d)
    eng.add_synthetic(31, 0, n, dist=0, normalize=True)
    eng.set_option("batch_mode", 0)
    x = O.synth_rows(31, 0, n, d)
    O.normalize_L2(x)
    adm = np.random.default_rng(5).random(n) < 0.5
    packed = np.zeros((n + 31) // 32 * 4, dtype=np.uint8)
    pk = mv.pack_mask(adm)
    packed[:pk.size] = pk
    m_dev = torch.from_numpy(packed.view(np.int32)).cuda()
    NQ = 48
    qh = O.synth_rows(32, 0, NQ, d)
    O.normalize_L2(qh)
    q = torch.from_numpy(qh).cuda()
    ws = eng.workspace()
    st = torch.cuda.current_stream().cuda_stream
    jobs = [(k, nq, i) for k in (10, 100) for nq in (1, 2, 4) for i in range(0, NQ, nq)]
    np.random.default_rng(9).shuffle(jobs)
    res = {}
    for pdl in (0, 1):
        eng.set_option("pdl", pdl)
        out = {(k, nq): (torch.empty(NQ, k, device="cuda"), torch.empty(NQ, k, dtype=torch.int64, device="cuda"))
               for k in (10, 100) for nq in (1, 2, 4)}
        for rep in range(3):
            for k, nq, i in jobs:
                D, I = out[(k, nq)]
                eng.search_device(ws, q[i:i + nq].data_ptr(), nq, k, D[i:i + nq].data_ptr(), I[i:i + nq].data_ptr(),
                                  m_dev.data_ptr(), n, stream=st)
        torch.cuda.synchronize()
        res[pdl] = {key: (D.cpu().numpy(), I.cpu().numpy()) for key, (D, I) in out.items()}
        # every search of one shape into ONE slot
        for k, nq in ((10, 1), (100, 2)):
            D1 = torch.empty(nq, k, device="cuda")
            I1 = torch.empty(nq, k, dtype=torch.int64, device="cuda")
            for i in range(0, NQ, nq):
                eng.search_device(ws, q[i:i + nq].data_ptr(), nq, k, D1.data_ptr(), I1.data_ptr(), m_dev.data_ptr(), n, stream=st)
            torch.cuda.synchronize()
            assert np.array_equal(I1.cpu().numpy(), res[pdl][(k, nq)][1][NQ - nq:])
            assert np.array_equal(D1.cpu().numpy(), res[pdl][(k, nq)][0][NQ - nq:])
    for key in res[0]:
        assert np.array_equal(res[0][key][1], res[1][key][1]) and np.array_equal(res[0][key][0], res[1][key][0]), key
    for k in (10, 100):
        Dr, Ir = O.search_masked(x, adm, qh, k)
        rep_ = O.classify_parity(x, qh, res[1][(k, 1)][1], res[1][(k, 1)][0], Ir, Dr, rel_tol=REL_TOL, admissible=adm)
        assert rep_["ok"], rep_
    eng.set_option("pdl", 0)
    del ws
    eng.close()


def test_coalesced_concurrent_searches_equal_direct_ones(mv):
    """Concurrent single-query calls are coalesced into shared passes (<= 8 per scan launch with
    per-query filters, tensor-core batch when unfiltered); every caller must get exactly what a
    lone call returns."""
    import threading
    n, d, k, nthreads, per = 60_000, 256, 10, 24, 12
    x, q = _data(n, d, nthreads * per, seed=12)
    eng = mv.FlatIPEngine(d)
    eng.add(x)
    eng.remove_rows(np.arange(0, n, 17))
    rng = np.random.default_rng(3)
    masks = [None if i % 3 == 0 else (rng.random(n) < (0.05 if i % 3 == 1 else 0.6)) for i in range(nthreads * per)]
    packed = [None if m is None else mv.pack_mask(m) for m in masks]
    eng.set_option("coalesce", 0)
    want = [eng.search(q[i:i + 1], k, mask=packed[i], mask_rows=n if packed[i] is not None else None)
            for i in range(nthreads * per)]
    eng.set_option("coalesce", 1)
    got = [None] * (nthreads * per)
    errs = []

    def worker(t):
        try:
            for j in range(per):
                i = t * per + j
                got[i] = eng.search(q[i:i + 1], k, mask=packed[i], mask_rows=n if packed[i] is not None else None)
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    ts = [threading.Thread(target=worker, args=(t,)) for t in range(nthreads)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errs, errs[:1]
    for i in range(nthreads * per):
        assert np.array_equal(got[i][1], want[i][1]), i
        assert np.array_equal(got[i][0], want[i][0]), i
    # mixed k and large k in flight at the same time
    def worker2(t):
        try:
            for j in range(6):
                kk = (3, 10, 200)[(t + j) % 3]
                D, I = eng.search(q[t:t + 1], kk)
                eng.set_option  # noqa: B018
                assert D.shape == (1, kk) and np.all(np.diff(D[0][I[0] >= 0]) <= 0)
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    ts = [threading.Thread(target=worker2, args=(t,)) for t in range(12)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errs, errs[:1]
    eng.close()


def test_mask_handles_and_coalesced_filtered_batches(mv):
    """Device-resident filters: same results as host-supplied masks, alone, in explicit
    batches (nq > 1, one common handle) and when many threads each bring their own handle
    (coalesced into a tensor-core batch with per-query filters)."""
    import threading
    n, d, k = 80_000, 384, 10
    x, q = _data(n, d, 96, seed=21)
    eng = mv.FlatIPEngine(d)
    eng.add(x[:70_000])
    rng = np.random.default_rng(4)
    adms = [rng.random(70_000) < f for f in (0.5, 0.1, 0.9, 0.001)]
    handles = [eng.mask_handle(a) for a in adms]
    eng.add(x[70_000:])           # rows appended after the handles were made are not admissible through them
    eng.remove_rows(np.arange(5, 70_000, 13))
    live = np.ones(n, dtype=bool)
    live[np.arange(5, 70_000, 13)] = False
    for a, h in zip(adms, handles):
        full = np.zeros(n, dtype=bool)
        full[:70_000] = a
        D, I = eng.search(q[:3], k, mask=h)            # nq > 1: common handle
        _check(x, q[:3], k, D, I, full & live)
        D1, I1 = eng.search(q[:1], k, mask=h)          # single query
        assert np.array_equal(I1[0], I[0]) and np.array_equal(D1[0], D[0])
        D40, I40 = eng.search(q[:40], k, mask=h)       # tensor-core batch with a common filter
        _check(x, q[:40], k, D40, I40, full & live)
    want = []
    for i in range(96):
        full = np.zeros(n, dtype=bool)
        full[:70_000] = adms[i % 4]
        want.append(O.search_masked(x, full & live, q[i:i + 1], k))
    got = [None] * 96
    errs = []

    def worker(t):
        try:
            for i in range(t, 96, 32):
                got[i] = eng.search(q[i:i + 1], k, mask=handles[i % 4] if i % 5 else None)
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    ts = [threading.Thread(target=worker, args=(t,)) for t in range(32)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errs, errs[:1]
    for i in range(96):
        if i % 5:
            full = np.zeros(n, dtype=bool)
            full[:70_000] = adms[i % 4]
            rep = O.classify_parity(x, q[i:i + 1], got[i][1], got[i][0], want[i][1], want[i][0], admissible=full & live)
        else:
            Dr, Ir = O.search_masked(x, live, q[i:i + 1], k)
            rep = O.classify_parity(x, q[i:i + 1], got[i][1], got[i][0], Ir, Dr, admissible=live)
        assert rep["ok"], (i, rep)
    [h.close() for h in handles]
    eng.close()


def test_children_may_outlive_their_index(mv):
    """Mask handles, columns and workspaces destroyed AFTER their index (what a garbage collector does) are inert,
    not dangling: no crash, no stale CUDA error leaking into the next call; a grouped index cannot go first."""
    import torch
    eng = mv.FlatIPEngine(64)
    eng.add_synthetic(1, 0, 5000, dist=0, normalize=True)
    adm = np.zeros(5000, dtype=bool)
    adm[::3] = True
    handle = eng.mask_handle(adm)
    col = eng.column()
    col.append(np.arange(5000, dtype=np.float64), np.ones(5000, dtype=np.uint8))
    pred = col.predicate("$gt", 10.0)
    ws = eng.workspace()
    q = O.synth_rows(2, 0, 1, 64)
    eng.search(q, 5, mask=handle)
    eng.close()                      # the index goes FIRST
    with pytest.raises(Exception):
        handle.count()
    for child in (handle, pred, col, ws):
        child.close()                # ... and its children afterwards: harmless
    eng2 = mv.FlatIPEngine(64)       # the next calls in this thread start from a clean error state
    eng2.add_synthetic(1, 0, 5000, dist=0, normalize=True)
    D, I = eng2.search(q, 5)
    assert I[0, 0] >= 0
    grp = mv.ShardGroup([eng2])
    with pytest.raises(Exception):
        eng2.close()                 # a member of a group cannot be destroyed before the group
    grp.close()
    eng2.close()


@pytest.mark.parametrize("n,d", [(20_000, 64), (100_000, 384), (60_000, 1000)])
def test_survivor_tail_is_result_neutral(mv, n, d):
    """32 < k <= 128: the survivor-list scan (shared threshold, one global list, last-CTA sort) returns exactly
    what the per-warp-select scan returns -- with filters, tombstones, duplicates (list overflow -> fallback),
    through the host API and the stream API."""
    import torch
    x, q = _data(n, d, 5, seed=71)
    eng = mv.FlatIPEngine(d)
    eng.set_option("coalesce", 0)
    eng.add(x)
    adm = np.random.default_rng(2).random(n) < 0.3
    for k in (17, 32, 33, 100, 128):
        eng.set_option("survivor_tail", 0)
        ref = [eng.search(q[i:i + 1], k) for i in range(5)] + [eng.search(q[i:i + 1], k, mask=adm) for i in range(5)]
        eng.set_option("survivor_tail", 1)
        for rep in range(2):
            got = [eng.search(q[i:i + 1], k) for i in range(5)] + [eng.search(q[i:i + 1], k, mask=adm) for i in range(5)]
            for (Dr, Ir), (Dg, Ig) in zip(ref, got):
                assert np.array_equal(Ir, Ig) and np.array_equal(Dr, Dg), (k, rep)
        _check(x, q[:1], k, *eng.search(q[:1], k))
    # a crowd of exact ties: 3000 copies of one row all tie at the top (the tail's selection falls back to a full sort) ...
    eng.add(np.repeat(x[3:4], 3000, axis=0))
    eng.set_option("survivor_tail", 0)
    Dr, Ir = eng.search(x[3:4], 100)
    eng.set_option("survivor_tail", 1)
    Dg, Ig = eng.search(x[3:4], 100)
    assert np.array_equal(Ir, Ig) and np.array_equal(Dr, Dg)
    # ... and overflow: 9000 copies -> the list overflows -> the classic scan answers
    eng.add(np.repeat(x[7:8], 9000, axis=0))
    eng.remove_rows(np.arange(0, 2000, 3))
    ws = eng.workspace()
    qd = torch.from_numpy(np.ascontiguousarray(x[7:8])).cuda()
    for k in (20, 100):
        eng.set_option("survivor_tail", 0)
        Dr, Ir = eng.search(x[7:8], k)
        eng.set_option("survivor_tail", 1)
        Dh, Ih = eng.search(x[7:8], k)                          # host API: re-run after the pinned flag
        assert np.array_equal(Ir, Ih) and np.array_equal(Dr, Dh)
        D = torch.empty(1, k, device="cuda")
        I = torch.empty(1, k, dtype=torch.int64, device="cuda")
        for _ in range(3):                                      # stream API: conditional launch; state resets itself
            eng.search_device(ws, qd.data_ptr(), 1, k, D.data_ptr(), I.data_ptr(), stream=torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            assert np.array_equal(I.cpu().numpy(), Ir) and np.array_equal(D.cpu().numpy(), Dr)
    ws.close()
    eng.close()


@pytest.mark.parametrize("n,d", [(3_000, 32), (150_000, 384), (400_000, 100)])
def test_host_path_pieces_are_result_neutral(mv, n, d):
    """Option "host_path": results written straight to pinned host memory, inputs pulled by a grid the scan depends
    on programmatically -- every combination returns what the copy-engine path
    returns, for unfiltered / per-call filter / resident handle searches, on the classic scan, the survivor-list scan
    (k = 100), the int8 shadow scan, query batches and large k (which stay on the copy path), with a different filter
    on every call (the staging buffer and the sequence word are reused) and from several host threads."""
    import threading
    x, q = _data(n, d, 12, seed=91)
    eng = mv.FlatIPEngine(d)
    eng.add(x)
    rng = np.random.default_rng(4)
    masks = [rng.random(n) < f for f in (0.5, 0.05, 0.9, 0.5, 0.001, 0.3)]
    handle = eng.mask_handle(masks[0])

    def sweep(shadow):
        eng.set_option("scan_shadow", shadow)
        out = []
        for k in (1, 10, 100):
            for i in range(6):
                out.append(eng.search(q[i:i + 1], k))
                out.append(eng.search(q[i:i + 1], k, mask=masks[i]))
                out.append(eng.search(q[i:i + 1], k, mask=masks[i][:n - 37 * i]))   # mask_rows < ntotal
                out.append(eng.search(q[i:i + 1], k, mask=handle))
        out.append(eng.search(q[:5], 10, mask=masks[1]))     # batch: copy path
        out.append(eng.search(q[:1], 300, mask=masks[2]))    # large k: copy path
        return out

    eng.set_option("coalesce", 0)
    for shadow in (0, 1):
        eng.set_option("host_path", 0)
        ref = sweep(shadow)
        for hp in (1, 2, 3):
            eng.set_option("host_path", hp)
            for rep in range(2):
                got = sweep(shadow)
                for j, ((Dr, Ir), (Dg, Ig)) in enumerate(zip(ref, got)):
                    assert np.array_equal(Ir, Ig) and np.array_equal(Dr, Dg), (shadow, hp, rep, j)
    eng.set_option("scan_shadow", 0)
    _check(x, q[:1], 10, *eng.search(q[:1], 10))
    _check(x, q[1:2], 10, *eng.search(q[1:2], 10, mask=masks[0]), adm=masks[0])
    # a filter the caller keeps in pinned memory is pulled from where it lies (whole 16-byte vectors; the last partial
    # one and a last byte with spare bits still go through the staging buffer)
    import torch
    eng.set_option("host_path", 3)
    for rows in (n, n - 3, n - 64, n - 129):
        pm = mv.pack_mask(masks[0][:rows])
        tp = torch.from_numpy(pm).pin_memory()
        for k in (10, 100):
            Dr, Ir = eng.search(q[:1], k, mask=pm, mask_rows=rows)
            Dg, Ig = eng.search(q[:1], k, mask=tp.numpy(), mask_rows=rows)
            assert np.array_equal(Ir, Ig) and np.array_equal(Dr, Dg), (rows, k)
            if len(pm) > 17:   # misaligned view of pinned memory: staged like any other buffer
                off = np.concatenate([np.zeros(1, np.uint8), pm])
                to = torch.from_numpy(off).pin_memory()
                Dg, Ig = eng.search(q[:1], k, mask=to.numpy()[1:], mask_rows=rows)
                assert np.array_equal(Ir, Ig) and np.array_equal(Dr, Dg), (rows, k, "misaligned")
    # several host threads, each with its own filter, default host path, coalescer on and off
    eng.set_option("host_path", 0)
    want = [eng.search(q[i:i + 1], 10, mask=masks[i % 6]) for i in range(12)]
    eng.set_option("host_path", 3)
    for co in (0, 1):
        eng.set_option("coalesce", co)
        errs = []

        def worker(t):
            try:
                for rep in range(40):
                    i = (t + rep) % 12
                    Dg, Ig = eng.search(q[i:i + 1], 10, mask=masks[i % 6])
                    assert np.array_equal(Ig, want[i][1]) and np.array_equal(Dg, want[i][0]), (co, t, rep)
            except Exception as e:   # noqa: BLE001
                errs.append(e)
        ts = [threading.Thread(target=worker, args=(t,)) for t in range(6)]
        [t.start() for t in ts]
        [t.join() for t in ts]
        assert not errs, errs[:1]
    handle.close()
    eng.close()


@pytest.mark.parametrize("n,d", [(6_000, 64), (250_000, 128)])
def test_large_k_fast_select_equals_the_radix_select(mv, n, d):
    """Host-buffer searches with 128 < k <= 8192 take the histogram select (two 12-bit digit passes, collect, one
    sort); it must return exactly what the radix select returns -- with filters (incl. fewer admissible rows than k),
    tombstones, several queries per call, k beyond its range, and a crowd of exact ties around the k-th score that
    overflows its list (the host then re-runs the query on the radix select)."""
    x, q = _data(n, d, 3, seed=17)
    eng = mv.FlatIPEngine(d)
    eng.set_option("coalesce", 0)
    eng.add(x)
    rng = np.random.default_rng(3)
    masks = [None, rng.random(n) < 0.5, rng.random(n) < 0.01]
    eng.remove_rows(np.arange(5, n, 11))

    def sweep():
        out = []
        for k in (129, 500, 1000, 4097, 8192, 9000):
            if k > n:
                continue
            for m in masks:
                out.append(eng.search(q[:1], k, mask=m))
            out.append(eng.search(q, k, mask=masks[1]))
        return out

    eng.set_option("large_k_fast", 0)
    ref = sweep()
    eng.set_option("large_k_fast", 1)
    for rep in range(2):
        for j, ((Dr, Ir), (Dg, Ig)) in enumerate(zip(ref, sweep())):
            assert np.array_equal(Ir, Ig) and np.array_equal(Dr, Dg), (rep, j)
    live = np.ones(n, bool); live[5::11] = False
    _check(x, q[:1], 500, *eng.search(q[:1], 500), adm=live)
    # 20 000 copies of one row: every one of them shares the k-th row's prefix -> list overflow -> radix select
    eng.add(np.repeat(x[3:4], 20_000, axis=0))
    for k in (200, 3000):
        eng.set_option("large_k_fast", 0)
        Dr, Ir = eng.search(x[3:4], k)
        eng.set_option("large_k_fast", 1)
        for _ in range(2):
            Dg, Ig = eng.search(x[3:4], k)
            assert np.array_equal(Ir, Ig) and np.array_equal(Dr, Dg), k
        Dg, Ig = eng.search(q[:1], k)        # and the next ordinary query finds the state clean
        eng.set_option("large_k_fast", 0)
        Dr, Ir = eng.search(q[:1], k)
        assert np.array_equal(Ir, Ig) and np.array_equal(Dr, Dg), k
    eng.close()
