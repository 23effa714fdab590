"""GPU: the opt-in int8 shadow mode ("scan_shadow" = 1) must return EXACTLY what the fp32 scan returns --
ids and distances bit-identical -- because the int8 pass only selects a candidate superset (rigorous
Cauchy-Schwarz bound) and the survivors are re-scored from the fp32 rows with the scan's summation order."""
import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mv():
    import minivectordb_b200 as m
    return m


def _both(eng, q, k, **kw):
    eng.set_option("scan_shadow", 0)
    ref = [eng.search(q[i:i + 1], k, **kw) for i in range(q.shape[0])]
    eng.set_option("scan_shadow", 1)
    got = [eng.search(q[i:i + 1], k, **kw) for i in range(q.shape[0])]
    for i, ((Dr, Ir), (Dg, Ig)) in enumerate(zip(ref, got)):
        assert np.array_equal(Ir, Ig), (i, Ir[0][:5], Ig[0][:5])
        assert np.array_equal(Dr, Dg), (i, Dr[0][:5], Dg[0][:5])
    return ref


@pytest.mark.parametrize("n,d", [(20_000, 64), (50_000, 100), (200_000, 384), (100_000, 512), (60_000, 768), (40_000, 1024),
                                 (33_333, 333)])
def test_shadow_mode_is_bit_identical_to_the_fp32_scan(mv, n, d):
    eng = mv.FlatIPEngine(d)
    eng.add_synthetic(1234, 0, n, dist=0, normalize=True)
    q = O.synth_rows(4321, 0, 6, d)
    O.normalize_L2(q)
    eng.set_option("coalesce", 0)
    for k in (1, 10, 100, 128):
        _both(eng, q, k)
    adm = np.random.default_rng(1).random(n) < 0.3
    _both(eng, q, 10, mask=adm)
    _both(eng, q[:2] * 3.0, 10, normalize=True)      # engine-side query normalisation
    _both(eng, q[:2] * 3.0, 10)                      # un-normalised queries
    # oracle parity of the shadow mode on its own
    x = eng.reconstruct_n(0, n)
    eng.set_option("scan_shadow", 1)
    D, I = eng.search(q[:1], 10)
    Dr, Ir = O.search_flat_ip(x, q[:1], 10)
    assert O.classify_parity(x, q[:1], I, D, Ir, Dr)["ok"]
    eng.close()


def test_shadow_mode_with_tombstones_appends_and_compaction(mv):
    n, d = 120_000, 256
    eng = mv.FlatIPEngine(d)
    eng.set_option("coalesce", 0)
    eng.add_synthetic(7, 0, n, dist=0, normalize=True)
    q = O.synth_rows(8, 0, 4, d)
    O.normalize_L2(q)
    ref0 = _both(eng, q, 10)
    top = np.unique(np.concatenate([r[1][0][:3] for r in ref0]))
    eng.remove_rows(top)                              # the best rows disappear
    ref1 = _both(eng, q, 10)
    assert not (set(np.concatenate([r[1][0] for r in ref1]).tolist()) & set(top.tolist()))
    eng.add_synthetic(7, n, 30_000, dist=0, normalize=True)   # the shadow follows appends
    _both(eng, q, 10)
    dup = eng.reconstruct_n(5, 1)
    eng.add(np.repeat(dup, 40, axis=0))               # exact duplicates: ties by ascending row, as the scan
    _both(eng, dup, 10)
    eng.compact()
    _both(eng, q, 10)
    eng.close()


def test_shadow_mode_uniform_positive_vectors_and_clustered_data(mv):
    """All-positive vectors (what the reference's tests use: np.random.rand) put every score in a narrow
    band -- the candidate set is large but the answer must not change; clustered data (many near-duplicates
    of a few centres) may overflow the candidate list, which must fall back to the fp32 scan on the device."""
    d = 384
    eng = mv.FlatIPEngine(d)
    eng.set_option("coalesce", 0)
    eng.add_synthetic(11, 0, 150_000, dist=1, normalize=True)
    q = O.synth_rows(12, 0, 4, d, O.DIST_UNIFORM)
    O.normalize_L2(q)
    _both(eng, q, 10)
    eng.close()
    rng = np.random.default_rng(5)
    centres = rng.standard_normal((4, d)).astype(np.float32)
    x = (centres[rng.integers(0, 4, 100_000)] + 1e-4 * rng.standard_normal((100_000, d))).astype(np.float32)
    eng = mv.FlatIPEngine(d)
    eng.set_option("coalesce", 0)
    eng.add(x, normalize=True)
    _both(eng, centres + 1e-3 * rng.standard_normal((4, d)).astype(np.float32), 10, normalize=True)
    eng.close()


def test_shadow_overflow_falls_back_on_the_device_and_on_the_host_path(mv):
    """Thousands of rows within the int8 error band of the k-th best (near-duplicates of the query's
    neighbourhood) overflow the survivor list: the stream-only API answers through the conditional fp32
    launch, the host-buffer API re-runs the query -- both return the fp32 scan's answer."""
    import torch
    d, n = 256, 60_000
    rng = np.random.default_rng(9)
    centre = rng.standard_normal(d).astype(np.float32)
    x = (centre[None, :] + 2e-4 * rng.standard_normal((n, d))).astype(np.float32)
    eng = mv.FlatIPEngine(d)
    eng.set_option("coalesce", 0)
    eng.add(x, normalize=True)
    q = (centre + 1e-3 * rng.standard_normal(d)).astype(np.float32)[None, :]
    q /= np.linalg.norm(q)
    eng.set_option("scan_shadow", 0)
    Dr, Ir = eng.search(q, 10)
    eng.set_option("scan_shadow", 1)
    Dh, Ih = eng.search(q, 10)                                  # host-buffer API
    assert np.array_equal(Ih, Ir) and np.array_equal(Dh, Dr)
    ws = eng.workspace()
    qd = torch.from_numpy(q).cuda()
    D = torch.empty(1, 10, device="cuda")
    I = torch.empty(1, 10, dtype=torch.int64, device="cuda")
    for _ in range(3):                                          # stream-only API, repeated: the state resets itself
        eng.search_device(ws, qd.data_ptr(), 1, 10, D.data_ptr(), I.data_ptr(), stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        cand, surv, overflowed = ws.shadow_counters()
        assert overflowed and surv > 4096, (cand, surv, overflowed)
        assert np.array_equal(I.cpu().numpy(), Ir) and np.array_equal(D.cpu().numpy(), Dr)
    # ordinary data on the same workspace afterwards: no overflow, few candidates
    eng2 = mv.FlatIPEngine(d)
    eng2.set_option("scan_shadow", 1)
    eng2.add_synthetic(3, 0, 200_000, dist=0, normalize=True)
    ws2 = eng2.workspace()
    q2 = O.synth_rows(4, 0, 1, d)
    O.normalize_L2(q2)
    q2d = torch.from_numpy(q2).cuda()
    eng2.search_device(ws2, q2d.data_ptr(), 1, 10, D.data_ptr(), I.data_ptr(), stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    cand, surv, overflowed = ws2.shadow_counters()
    assert not overflowed and 10 <= surv <= cand < 5000, (cand, surv)
    eng2.set_option("scan_shadow", 0)
    Dr2, Ir2 = eng2.search(q2, 10)
    assert np.array_equal(I.cpu().numpy(), Ir2) and np.array_equal(D.cpu().numpy(), Dr2)
    ws.close(); ws2.close(); eng.close(); eng2.close()


@pytest.mark.parametrize("coalesce", [0, 1])
def test_shadow_mode_from_many_host_threads(mv, coalesce):
    """Concurrent single-query callers in shadow mode: every call runs on its own pooled workspace (own int8
    state); with the coalescer on, lone calls take the shadow pass and shared passes take the batch paths --
    every caller still gets the fp32 scan's answer."""
    import threading
    n, d, k = 150_000, 256, 10
    eng = mv.FlatIPEngine(d)
    eng.add_synthetic(61, 0, n, dist=0, normalize=True)
    q = O.synth_rows(62, 0, 16, d)
    O.normalize_L2(q)
    adm = np.random.default_rng(6).random(n) < 0.5
    eng.set_option("coalesce", 0)
    eng.set_option("scan_shadow", 0)
    ref = [eng.search(q[i:i + 1], k, mask=adm if i % 2 else None) for i in range(16)]
    eng.set_option("scan_shadow", 1)
    eng.set_option("coalesce", coalesce)
    errs = []

    def run(i):
        try:
            for _ in range(25):
                D, I = eng.search(q[i:i + 1], k, mask=adm if i % 2 else None)
                assert np.array_equal(I, ref[i][1]) and np.array_equal(D, ref[i][0])
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    ts = [threading.Thread(target=run, args=(i,)) for i in range(16)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    eng.close()
    assert not errs, errs[:1]
