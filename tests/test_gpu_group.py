"""GPU: the shard group (mvdb_group_*) -- one process, several shards, fused NVLink exchange.
Shards are dealt over the visible GPUs; on a one-GPU box they share the device and the exchange
runs between kernels on different streams (same code path, peer == self)."""
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _setup(nshards, n, d, seed=21, empty=()):
    from minivectordb_b200 import FlatIPEngine, ShardGroup, _native
    from oracle import oracle as O
    ndev = _native.device_count()
    x = O.synth_rows(seed, 0, n, d)
    O.normalize_L2(x)
    live = [s for s in range(nshards) if s not in empty]
    cuts = np.linspace(0, n, len(live) + 1).astype(int)
    cuts[1:-1] += 13   # uneven
    bounds = {}
    for i, s in enumerate(live):
        bounds[s] = (int(cuts[i]), int(cuts[i + 1]))
    engines = []
    for s in range(nshards):
        e = FlatIPEngine(d, device=s % ndev)
        if s in bounds:
            e.add(x[bounds[s][0]:bounds[s][1]])
        engines.append(e)
    return x, bounds, engines, ShardGroup(engines)


def _globalise(S, R, bounds):
    I = np.full(S.shape, -1, dtype=np.int64)
    for s, (lo, hi) in bounds.items():
        sel = S == s
        I[sel] = R[sel] + lo
    return I


@pytest.mark.parametrize("nshards,empty", [(1, ()), (2, ()), (3, (1,)), (8, (0, 5))])
def test_group_matches_oracle(nshards, empty):
    from oracle import oracle as O
    n, d = 30000, 384
    x, bounds, engines, grp = _setup(nshards, n, d, empty=empty)
    try:
        for nq in (1, 3, 11):
            q = O.synth_rows(22 + nq, 0, nq, d)
            O.normalize_L2(q)
            for k in (1, 10, 128):
                D, S, R = grp.search(q, k)
                I = _globalise(S, R, bounds)
                Dr, Ir = O.search_flat_ip(x, q, k)
                rep = O.classify_parity(x, q, I, D, Ir, Dr)
                assert rep["ok"], (nshards, nq, k, rep)
        # filters: bool arrays, resident handles, and a mix (None = that shard unfiltered)
        q = O.synth_rows(29, 0, 4, d)
        O.normalize_L2(q)
        adm = np.random.default_rng(3).random(n) < 0.3
        Dr, Ir = O.search_masked(x, adm, q, 10)
        as_bool = [adm[bounds[s][0]:bounds[s][1]] if s in bounds else np.zeros(0, dtype=bool) for s in range(nshards)]
        as_handle = [engines[s].mask_handle(m) if s in bounds else m for s, m in enumerate(as_bool)]
        mixed = [as_handle[s] if s % 2 else as_bool[s] for s in range(nshards)]
        for masks in (as_bool, as_handle, mixed):
            for _ in range(6):   # sequence numbers / parity double buffering
                D, S, R = grp.search(q, 10, masks=masks)
            rep = O.classify_parity(x, q, _globalise(S, R, bounds), D, Ir, Dr, admissible=adm)
            assert rep["ok"], (nshards, rep)
        first = min(bounds)
        only = [None if s == first else np.zeros(bounds[s][1] - bounds[s][0] if s in bounds else 0, dtype=bool)
                for s in range(nshards)]
        D, S, R = grp.search(q, 10, masks=only)
        assert (S == first).all()
        none = [np.zeros(bounds[s][1] - bounds[s][0] if s in bounds else 0, dtype=bool) for s in range(nshards)]
        D, S, R = grp.search(q, 10, masks=none)
        assert (S == -1).all() and (R == -1).all() and (D == np.finfo(np.float32).min).all()
        with pytest.raises(Exception):
            grp.search(q, 129)
    finally:
        grp.close()
        for e in engines:
            e.close()


def test_group_sees_inserts_and_deletes_of_its_shards():
    from oracle import oracle as O
    n, d = 8000, 128
    x, bounds, engines, grp = _setup(3, n, d, seed=31)
    try:
        q = x[[100, 4000, 7900]].copy()
        D, S, R = grp.search(q, 1)
        assert _globalise(S, R, bounds)[:, 0].tolist() == [100, 4000, 7900]
        s, (lo, hi) = 1, bounds[1]
        victim = 4000 - lo
        engines[1].remove_rows([victim])
        D, S, R = grp.search(q[1:2], 1)
        assert not (S[0, 0] == 1 and R[0, 0] == victim)
        extra = O.synth_rows(32, 0, 5, d)
        O.normalize_L2(extra)
        first = engines[2].add(extra)
        D, S, R = grp.search(extra[3:4], 1)
        assert S[0, 0] == 2 and R[0, 0] == first + 3 and abs(D[0, 0] - 1.0) < 1e-5
    finally:
        grp.close()
        for e in engines:
            e.close()


def test_group_search_from_many_threads():
    import threading
    from oracle import oracle as O
    n, d = 20000, 256
    x, bounds, engines, grp = _setup(4, n, d, seed=41)
    q = O.synth_rows(42, 0, 16, d)
    O.normalize_L2(q)
    Dr, Ir = O.search_flat_ip(x, q, 10)
    errs = []

    def run(i):
        try:
            for _ in range(20):
                D, S, R = grp.search(q[i:i + 1], 10)
                rep = O.classify_parity(x, q[i:i + 1], _globalise(S, R, bounds), D, Ir[i:i + 1], Dr[i:i + 1])
                assert rep["ok"], rep
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    ts = [threading.Thread(target=run, args=(i,)) for i in range(8)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    grp.close()
    for e in engines:
        e.close()
    assert not errs, errs[:1]


def test_group_with_the_int8_shadow_mode_is_bit_identical():
    """Shards in "scan_shadow" mode: the int8 pass + exact re-scoring runs on every device and ITS tail does the
    NVLink exchange -- the merged answer is bit-identical to the fp32 group search."""
    from oracle import oracle as O
    n, d = 120_000, 384
    x, bounds, engines, grp = _setup(3, n, d, seed=51)
    try:
        q = O.synth_rows(52, 0, 6, d)
        O.normalize_L2(q)
        adm = np.random.default_rng(4).random(n) < 0.4
        masks = [adm[bounds[s][0]:bounds[s][1]] for s in range(3)]
        for k in (1, 10, 100):
            ref = [grp.search(q[i:i + 1], k) for i in range(6)]
            refm = [grp.search(q[i:i + 1], k, masks=masks) for i in range(6)]
            for e in engines:
                e.set_option("scan_shadow", 1)
            for rep in range(2):
                got = [grp.search(q[i:i + 1], k) for i in range(6)]
                gotm = [grp.search(q[i:i + 1], k, masks=masks) for i in range(6)]
                for a, b in zip(ref + refm, got + gotm):
                    assert all(np.array_equal(u, v) for u, v in zip(a, b)), k
            for e in engines:
                e.set_option("scan_shadow", 0)
    finally:
        grp.close()
        for e in engines:
            e.close()
