"""GPU: the tcgen05/TMEM batched path (large query batches, BASELINE config 3 shape).

* the raw tensor-core GEMM against an independent bf16 matmul (torch);
* "exact" mode (bf16 candidates + fp32 re-scoring) must return BIT-IDENTICAL
  ids and distances to the fp32 scan;
* "bf16" mode: recall@k against the fp32 oracle."""
import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mv():
    import minivectordb_b200 as m
    return m


def _data(n, d, nq, seed=1):
    x = O.synth_rows(seed, 0, n, d)
    O.normalize_L2(x)
    q = O.synth_rows(seed + 100, 0, nq, d)
    O.normalize_L2(q)
    return x, q


@pytest.mark.parametrize("n,d,nq", [(256, 64, 1), (1000, 64, 128), (777, 384, 200), (5000, 1000, 130),
                                    (4096, 1024, 256), (300, 40, 5)])
def test_tcgen05_gemm_matches_bf16_matmul(mv, n, d, nq):
    import torch
    x, q = _data(n, d, nq, seed=n + d)
    eng = mv.FlatIPEngine(d)
    eng.set_option("batch_cost_model", 0)   # always exercise the tensor path here
    eng.add(x)
    got = eng.debug_gemm_scores(q)
    xb = torch.from_numpy(x).to(torch.bfloat16).to(torch.float64)
    qb = torch.from_numpy(q).to(torch.bfloat16).to(torch.float64)
    want = (qb @ xb.T).numpy()
    assert got.shape == want.shape
    err = np.abs(got - want).max()
    assert err < 2e-5, err   # products of bf16 are exact in fp32; only the accumulation order differs
    eng.close()


@pytest.mark.parametrize("n,d,nq,k", [(50_000, 384, 200, 10), (30_000, 1024, 129, 100), (9_000, 100, 64, 10),
                                      (2_000, 768, 33, 128)])
def test_batched_exact_is_bit_identical_to_the_scan(mv, n, d, nq, k):
    x, q = _data(n, d, nq, seed=k + d)
    eng = mv.FlatIPEngine(d)
    eng.set_option("batch_cost_model", 0)   # always exercise the tensor path here
    eng.add(x)
    rng = np.random.default_rng(0)
    dead = rng.choice(n, n // 10, replace=False)
    adm = rng.random(n) < 0.5
    for step in range(3):
        mask = adm if step >= 1 else None
        if step == 2:
            eng.remove_rows(dead)
        eng.set_option("batch_mode", 0)
        Ds, Is = eng.search(q, k, mask=mask)
        eng.set_option("batch_mode", 1)
        Db, Ib = eng.search(q, k, mask=mask)
        assert np.array_equal(Is, Ib), (step, np.argwhere(Is != Ib)[:5])
        assert np.array_equal(Ds, Db), step
    # and against the oracle
    live = np.ones(n, dtype=bool)
    live[dead] = False
    Dr, Ir = O.search_masked(x, adm & live, q, k)
    rep = O.classify_parity(x, q, Ib, Db, Ir, Dr, admissible=adm & live)
    assert rep["ok"], rep
    eng.close()


def test_batched_bf16_mode_recall(mv):
    n, d, nq, k = 100_000, 384, 256, 10
    x, q = _data(n, d, nq, seed=3)
    eng = mv.FlatIPEngine(d)
    eng.set_option("batch_cost_model", 0)   # always exercise the tensor path here
    eng.add(x)
    eng.set_option("batch_mode", 2)
    D, I = eng.search(q, k)
    Dr, Ir = O.search_flat_ip(x, q, k)
    recall = np.mean([len(set(I[i]) & set(Ir[i])) / k for i in range(nq)])
    assert recall > 0.9, recall
    assert np.all(np.diff(D, axis=1) <= 0)
    assert np.abs(D - Dr).max() < 1e-2
    eng.close()


@pytest.mark.parametrize("n,d,nq,k", [(100_000, 384, 256, 10), (40_000, 1000, 130, 100), (9_000, 72, 64, 10)])
def test_batched_tf32_mode(mv, n, d, nq, k):
    """batch_mode 3: tcgen05 kind::tf32 over the fp32 matrix itself (no bf16 shadow).  Operands are
    truncated to 10 mantissa bits, so |score - exact| <= 2 * 2^-10 * |q||x| (+ fp32 accumulation) --
    a systematic shrink of the score rather than noise, which is why recall stays above bf16's;
    recall against the fp32 oracle, with tombstones and a filter mask applied."""
    x, q = _data(n, d, nq, seed=n % 97)
    eng = mv.FlatIPEngine(d)
    eng.set_option("batch_cost_model", 0)
    eng.add(x)
    eng.set_option("batch_mode", 3)
    D, I = eng.search(q, k)
    Dr, Ir = O.search_flat_ip(x, q, k)
    recall = np.mean([len(set(I[i]) & set(Ir[i])) / k for i in range(nq)])
    assert recall > 0.97, recall
    assert np.all(np.diff(D, axis=1) <= 0)
    exact = np.einsum("qkd,qd->qk", x[I].astype(np.float64), q.astype(np.float64))
    assert np.abs(D - exact).max() < 2.2e-3          # of the rows it DID return
    # masks and tombstones
    eng.set_option("batch_mode", 3)
    eng.remove_rows(np.arange(5, n, 9))
    live = np.ones(n, dtype=bool)
    live[np.arange(5, n, 9)] = False
    adm = np.random.default_rng(4).random(n) < 0.4
    D, I = eng.search(q, k, mask=adm)
    assert np.all(I >= 0) and np.all((adm & live)[I])
    Dr, Ir = O.search_masked(x, adm & live, q, k)
    recall = np.mean([len(set(I[i]) & set(Ir[i])) / k for i in range(nq)])
    assert recall > 0.97, recall
    eng.close()


def test_batched_overflow_falls_back_to_the_scan(mv):
    """Rows ordered by INCREASING similarity to a query make every row beat the
    running threshold: the candidate list overflows and that query is redone by
    the exact scan."""
    n, d, nq, k = 30_000, 64, 40, 10
    x, q = _data(n, d, nq, seed=5)
    order = np.argsort(x @ q[0])
    x = np.ascontiguousarray(x[order])
    eng = mv.FlatIPEngine(d)
    eng.set_option("batch_cost_model", 0)   # always exercise the tensor path here
    eng.add(x)
    D, I = eng.search(q, k)
    Dr, Ir = O.search_flat_ip(x, q, k)
    rep = O.classify_parity(x, q, I, D, Ir, Dr)
    assert rep["ok"], rep
    eng.close()


def test_batched_path_follows_appends(mv):
    x, q = _data(20_000, 128, 64, seed=9)
    eng = mv.FlatIPEngine(128)
    eng.set_option("batch_cost_model", 0)
    eng.add(x[:5000])
    eng.search(q, 10)              # builds the bf16 shadow for 5000 rows
    eng.add(x[5000:])              # shadow must be extended lazily
    D, I = eng.search(q, 10)
    Dr, Ir = O.search_flat_ip(x, q, 10)
    rep = O.classify_parity(x, q, I, D, Ir, Dr)
    assert rep["ok"], rep
    eng.close()


def test_batched_is_robust_to_row_order_and_dead_prefixes(mv):
    """The threshold schedule samples rows at power-of-16 strides, so neither a deleted /
    filtered-out PREFIX (time-ordered data with a date filter, churn deleting the oldest
    rows) nor rows sorted by similarity may flood the candidate lists."""
    n, d, nq, k = 120_000, 128, 48, 10
    x, q = _data(n, d, nq, seed=31)
    eng = mv.FlatIPEngine(d)
    eng.set_option("batch_cost_model", 0)   # always exercise the tensor path here
    eng.add(x)
    eng.remove_rows(np.arange(0, 70_000))             # the first 58 % of the rows are tombstones
    live = np.ones(n, dtype=bool)
    live[:70_000] = False
    D, I = eng.search(q, k)
    Dr, Ir = O.search_masked(x, live, q, k)
    assert O.classify_parity(x, q, I, D, Ir, Dr, admissible=live)["ok"]
    late = np.zeros(n, dtype=bool)
    late[100_000:] = True                             # filter admits only the newest rows
    D, I = eng.search(q, k, mask=late)
    Dr, Ir = O.search_masked(x, late, q, k)
    assert O.classify_parity(x, q, I, D, Ir, Dr, admissible=late)["ok"]
    eng.set_option("batch_mode", 0)
    Ds, Is = eng.search(q, k, mask=late)
    assert np.array_equal(I, Is) and np.array_equal(D, Ds)
    eng.close()


def test_batched_adversarial_periodic_filter_overflows_and_falls_back(mv):
    """Only ODD rows admissible: every sampling level (even strides) sees nothing, the last
    level floods the candidate lists, and the queries are redone by the exact scan."""
    n, d, nq, k = 40_000, 64, 40, 10
    x, q = _data(n, d, nq, seed=33)
    eng = mv.FlatIPEngine(d)
    eng.set_option("batch_cost_model", 0)   # always exercise the tensor path here
    eng.add(x)
    odd = (np.arange(n) % 2) == 1
    D, I = eng.search(q, k, mask=odd)
    Dr, Ir = O.search_masked(x, odd, q, k)
    assert O.classify_parity(x, q, I, D, Ir, Dr, admissible=odd)["ok"]
    eng.close()


@pytest.mark.parametrize("variant", [1, 2, 3])
@pytest.mark.parametrize("n,d,nq", [(1000, 64, 256), (5000, 1000, 300), (4096, 1024, 512), (70_000, 384, 1000)])
def test_tcgen05_2cta_gemm_matches_bf16_matmul(mv, n, d, nq, variant):
    """cta_group::2 variant (CTA pairs, 256 x 256 tiles) and TMA-multicast variant (cluster of 2
    sharing the X tile): same raw scores as the bf16 matmul."""
    import torch
    x, q = _data(n, d, nq, seed=n + d + 1)
    eng = mv.FlatIPEngine(d)
    eng.set_option("batch_cost_model", 0)   # always exercise the tensor path here
    eng.set_option("gemm_variant", variant)
    eng.add(x)
    got = eng.debug_gemm_scores(q)
    xb = torch.from_numpy(x).to(torch.bfloat16).to(torch.float64)
    qb = torch.from_numpy(q).to(torch.bfloat16).to(torch.float64)
    want = (qb @ xb.T).numpy()
    err = np.abs(got - want).max()
    assert err < 2e-5, err
    eng.close()


def test_batched_2cta_exact_is_bit_identical_to_the_scan(mv):
    n, d, nq, k = 60_000, 384, 700, 10
    x, q = _data(n, d, nq, seed=77)
    eng = mv.FlatIPEngine(d)
    eng.set_option("batch_cost_model", 0)   # always exercise the tensor path here
    eng.add(x)
    adm = np.random.default_rng(0).random(n) < 0.5
    eng.set_option("batch_mode", 0)
    Ds, Is = eng.search(q[:64], k, mask=adm)
    eng.set_option("batch_mode", 1)
    eng.set_option("gemm_variant", 0)
    D0, I0 = eng.search(q, k, mask=adm)
    assert np.array_equal(Is, I0[:64]) and np.array_equal(Ds, D0[:64])
    for variant in (1, 2, 3):
        eng.set_option("gemm_variant", variant)
        Db, Ib = eng.search(q, k, mask=adm)
        assert np.array_equal(I0, Ib) and np.array_equal(D0, Db), variant
    eng.close()


def test_full_size_config3_properties(mv):
    """BASELINE config 3 at full size (10 M x 1024, 4096 queries, k = 100) through the tensor-core
    path: size-independent properties on the whole batch, and bit-identity with the fp32 scan on a
    sample of the queries (the CPU oracle cannot finish this size in seconds)."""
    import torch
    free_b, _ = torch.cuda.mem_get_info()
    if free_b < 80 * (1 << 30):
        pytest.skip("needs ~65 GB of free HBM")
    n, d, nq, k = 10_000_000, 1024, 4096, 100
    eng = mv.FlatIPEngine(d, capacity_hint=n)
    eng.add_synthetic(1234, 0, n, dist=0, normalize=True)
    q = O.synth_rows(4321, 0, nq, d)
    O.normalize_L2(q)
    D, I = eng.search(q, k)                                   # exact mode (default)
    assert np.all(np.diff(D, axis=1) <= 0)                    # best first
    assert I.min() >= 0 and I.max() < n
    assert all(len(set(row)) == k for row in I[::64])         # no duplicates
    D2, I2 = eng.search(q, k)
    assert np.array_equal(I, I2) and np.array_equal(D, D2)    # idempotent
    eng.set_option("batch_mode", 0)
    sample = np.arange(0, nq, 512)
    Ds, Is = eng.search(q[sample], k)                         # fp32 scan, 8 queries
    assert np.array_equal(Is, I[sample]) and np.array_equal(Ds, D[sample])
    rows = eng.reconstruct_n(int(I[0, 0]), 1)[0]              # score re-derivation from the stored row
    assert abs(float(rows.astype(np.float64) @ q[0].astype(np.float64)) - float(D[0, 0])) < 1e-5 * abs(float(D[0, 0]))
    eng.set_option("batch_mode", 2)                           # bf16 mode: recall against exact
    Db, Ib = eng.search(q[:256], k)
    recall = np.mean([len(set(Ib[i]) & set(I[i])) / k for i in range(256)])
    assert recall > 0.98, recall
    eng.close()
