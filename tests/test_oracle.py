"""CPU: the oracle (C restatement of faiss flat-IP) against its float64 gold and
against the behaviours the reference's tests pin at the faiss boundary."""
import numpy as np
import pytest

from oracle import oracle as O


def _data(n, d, nq, seed=1):
    x = O.synth_rows(seed, 0, n, d)
    O.normalize_L2(x)
    q = O.synth_rows(seed + 100, 0, nq, d)
    O.normalize_L2(q)
    return x, q


def test_synth_c_matches_numpy_bit_for_bit():
    for dist in (O.DIST_BELL, O.DIST_UNIFORM):
        a = O.synth_rows(77, 12345, 257, 33, dist)
        b = O.synth_rows_numpy(77, 12345, 257, 33, dist)
        assert np.array_equal(a, b)
    u = O.synth_rows(5, 0, 1000, 16, O.DIST_UNIFORM)
    assert u.min() >= 0.0 and u.max() < 1.0


def test_normalize_L2_semantics():
    x = O.synth_rows(3, 0, 50, 19)
    x[7] = 0.0  # zero rows stay untouched (faiss: only if nr > 0)
    ref = x.astype(np.float64)
    nr = np.sqrt((ref ** 2).sum(1, keepdims=True))
    nr[nr == 0] = 1.0
    ref = ref / nr
    O.normalize_L2(x)
    assert np.allclose(x, ref, rtol=0, atol=3e-7)
    assert np.all(x[7] == 0.0)
    with pytest.raises(TypeError):
        O.normalize_L2(x.astype(np.float64))


@pytest.mark.parametrize("k", [1, 5, 10, 99, 100, 250])
def test_search_matches_gold(k):
    x, q = _data(5000, 64, 6)
    D, I = O.search_flat_ip(x, q, k)
    Dg, Ig = O.gold_topk(x, q, k)
    rep = O.classify_parity(x, q, I, D, Ig, Dg)
    assert rep["ok"], rep
    assert np.all(np.diff(D, axis=1) <= 0)  # best first


def test_k_larger_than_n_pads_like_faiss():
    x, q = _data(5, 8, 2)
    D, I = O.search_flat_ip(x, q, 8)
    assert np.all(I[:, 5:] == -1)
    assert np.all(D[:, 5:] == np.finfo(np.float32).min)
    assert sorted(I[0, :5].tolist()) == [0, 1, 2, 3, 4]


def test_gathered_branch_equals_masked_gold():
    x, q = _data(3000, 32, 4)
    adm = np.random.default_rng(0).random(3000) < 0.3
    D, I = O.search_masked(x, adm, q, 10)
    Dg, Ig = O.gold_topk(x, q, 10, adm)
    rep = O.classify_parity(x, q, I, D, Ig, Dg, admissible=adm)
    assert rep["ok"], rep
    assert adm[I[I >= 0]].all()


def test_gathered_respects_given_row_order():
    # the reference gathers rows in Python-set iteration order (VDB:510); positions
    # returned are into THAT list
    x, q = _data(100, 16, 1)
    rows = np.array([50, 3, 77, 10], dtype=np.int64)
    D, P = O.search_gathered(x, rows, q, 4)
    scores = x[rows] @ q[0]
    assert P[0].tolist() == np.argsort(-scores, kind="stable").tolist()


def test_duplicates_are_exact_ties():
    x, q = _data(10, 16, 1)
    xd = np.repeat(x[:1], 6, axis=0)
    D, I = O.search_flat_ip(xd, q, 4)
    assert len(set(D[0].tolist())) == 1
    assert len(set(I[0].tolist())) == 4


def test_faiss_shaped_index_object():
    x, q = _data(200, 12, 3)
    idx = O.IndexFlatIP(12)
    idx.add(x[:120])
    idx.add(x[120:])
    assert idx.ntotal == 200
    D, I = idx.search(q, 7)
    D2, I2 = O.search_flat_ip(x, q, 7)
    assert np.array_equal(I, I2) and np.array_equal(D, D2)


def test_reservoir_and_blocked_sgemm_paths_agree_with_gold():
    """k >= 100 runs faiss's ReservoirTopN (capacity (2k+15)&~15, partition_fuzzy); nq >= 20 runs the
    1024-row sgemm blocks feeding the block result handlers.  Both must select what the float64 gold
    selects (near-ties excused), on random rows and on heavy duplicates (exact ties)."""
    import numpy as np
    from oracle import oracle as O
    n, d = 20000, 64
    x = O.synth_rows(5, 0, n, d)
    O.normalize_L2(x)
    q = O.synth_rows(6, 0, 24, d)
    O.normalize_L2(q)
    assert O._lib().orc_reservoir_capacity(100) == 208 and O._lib().orc_reservoir_capacity(128) == 256
    for k in (1, 10, 99, 100, 128, 500):
        Dg, Ig = O.gold_topk(x, q, k)
        for D, I in (O.search_flat_ip(x, q, k), O.search_flat_ip_blas(x, q, k)):
            rep = O.classify_parity(x, q, I, D, Ig, Dg)
            assert rep["ok"], (k, rep)
            assert np.all(np.diff(D, axis=1) <= 0)
    xd = np.repeat(x[:300], 30, axis=0).copy()
    for k in (10, 100, 130):
        Dg, Ig = O.gold_topk(xd, q[:4], k)
        for D, I in (O.search_flat_ip(xd, q[:4], k), O.search_flat_ip_blas(xd, q[:4], k)):
            rep = O.classify_parity(xd, q[:4], I, D, Ig, Dg)
            assert rep["ok"] and rep["real_error"] == 0, (k, rep)
            assert all(len(set(r.tolist())) == k for r in I)          # no id twice
    # streamed matrix = resident matrix
    def chunk(r0, m):
        return x[r0:r0 + m]
    Db, Ib = O.search_flat_ip_blas(None, q, 100, make_chunk=chunk, n=n)
    Dr, Ir = O.search_flat_ip_blas(x, q, 100)
    assert np.array_equal(Ib, Ir) and np.array_equal(Db, Dr)
    # fewer rows than k: padding
    D, I = O.search_flat_ip_blas(x[:50], q, 100)
    assert (I[:, 50:] == -1).all() and (I[:, :50] >= 0).all() and (D[:, 50:] == O.FLT_LOWEST).all()
