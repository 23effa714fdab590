"""GPU: row-sharded search, one process per GPU (world = min(8, visible GPUs)): fused NVLink exchange
and NCCL transport against the oracle, with the final kernels.  With a single visible GPU the same
code path runs at world 1 (the exchange kernels are then covered by tests/test_gpu_group.py, which
puts several shards on one device)."""
import os
import sys
import time

import numpy as np
import pytest
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _words(loc):
    import minivectordb_b200 as mv
    words = np.zeros((loc.size + 31) // 32 * 4, dtype=np.uint8)
    pk = mv.pack_mask(loc)
    words[:pk.size] = pk
    return words


def _run_transport(rank, world, transport):
    from minivectordb_b200.distributed import RowShardedIndex
    from oracle import oracle as O
    n, d, k = 40000, 384, 10
    x = O.synth_rows(11, 0, n, d)
    O.normalize_L2(x)
    q = O.synth_rows(12, 0, 11, d)
    O.normalize_L2(q)
    # uneven shards; with >= 3 ranks rank 1 holds NOTHING (it still takes part in every exchange)
    holders = [r for r in range(world) if not (world >= 3 and r == 1)]
    cuts = np.linspace(0, n, len(holders) + 1).astype(int)
    cuts[1:-1] += 37
    bounds = np.zeros(world + 1, dtype=int)
    for r in range(world):
        bounds[r + 1] = cuts[holders.index(r) + 1] if r in holders else bounds[r]
    idx = RowShardedIndex(d, device=rank, exchange=transport)
    assert world == 1 or idx.exchange == transport
    idx.add(x[bounds[rank]:bounds[rank + 1]], normalize=False)
    assert idx.offset == bounds[rank] and idx.ntotal_global == n
    adm = np.random.default_rng(0).random(n) < 0.3
    loc = adm[bounds[rank]:bounds[rank + 1]]
    refs = {}
    for pdl in (0, 1):
        idx.engine.set_option("pdl", pdl)
        for kk in (1, k, 128):
            for nq in (1, 3, 8, 11):
                D, I = idx.search(q[:nq], kk)
                if (kk, nq, 0) not in refs:
                    refs[(kk, nq, 0)] = O.search_flat_ip(x, q[:nq], kk)
                Dr, Ir = refs[(kk, nq, 0)]
                rep = O.classify_parity(x, q[:nq], I, D, Ir, Dr)
                assert rep["ok"], (transport, pdl, kk, nq, rep)
                D, I = idx.search(q[:nq], kk, mask_local=loc)
                if (kk, nq, 1) not in refs:
                    refs[(kk, nq, 1)] = O.search_masked(x, adm, q[:nq], kk)
                Dr, Ir = refs[(kk, nq, 1)]
                rep = O.classify_parity(x, q[:nq], I, D, Ir, Dr, admissible=adm)
                assert rep["ok"], (transport, pdl, kk, nq, rep)
        # many rounds back to back: sequence numbers / parity double-buffering, host-buffer fast path
        Dm, Im = idx.search(q[:4], k, mask_local=loc)
        words = _words(loc)
        for i in range(24):
            Dp, Ip = idx.search_packed(q[i % 4:i % 4 + 1], k, words, loc.size)
            assert np.array_equal(Ip[0], Im[i % 4]) and np.array_equal(Dp[0], Dm[i % 4]), (transport, pdl, i)
    idx.engine.set_option("pdl", 0)
    # a filter that admits rows of the first shard only
    only0 = np.zeros(n, dtype=bool)
    only0[:50] = True
    D, I = idx.search(q, k, mask_local=only0[bounds[rank]:bounds[rank + 1]])
    Dr, Ir = O.search_masked(x, only0, q, k)
    assert O.classify_parity(x, q, I, D, Ir, Dr, admissible=only0)["ok"], transport
    assert not idx.exchange_timed_out()
    idx.close()


def _run_shadow_mode(rank, world):
    """int8 shadow mode under the fused exchange: every rank's int8 tail sends / merges; answers are
    bit-identical to the fp32 sharded search."""
    from minivectordb_b200.distributed import RowShardedIndex
    from oracle import oracle as O
    d, per = 256, 30_000
    n = per * world
    idx = RowShardedIndex(d, device=rank, exchange="fused" if world > 1 else "auto")
    idx.add(synthetic=(77, rank * per, per, 0), normalize=True)
    q = O.synth_rows(78, 0, 5, d)
    O.normalize_L2(q)
    ref = [idx.search(q[i:i + 1], 10) for i in range(5)]
    idx.engine.set_option("scan_shadow", 1)
    for rep in range(3):
        for i in range(5):
            D, I = idx.search(q[i:i + 1], 10)
            assert np.array_equal(I, ref[i][1]) and np.array_equal(D, ref[i][0]), (rep, i)
    x = O.synth_rows(77, 0, n, d)
    O.normalize_L2(x)
    Dr, Ir = O.search_flat_ip(x, q, 10)
    D = np.concatenate([r[0] for r in ref]); I = np.concatenate([r[1] for r in ref])
    assert O.classify_parity(x, q, I, D, Ir, Dr)["ok"]
    idx.close()


def _run_stable_numbering(rank, world):
    from minivectordb_b200.distributed import RowShardedIndex
    from oracle import oracle as O
    d = 128
    x = O.synth_rows(51, 0, 3000, d)
    O.normalize_L2(x)
    st = RowShardedIndex(d, device=rank, numbering="stable")
    labels = st.add_balanced(x, normalize=False)      # every rank passes the same block
    owners = labels >> RowShardedIndex.STABLE_SHIFT
    assert max(st.counts) - min(st.counts) <= 1 and sorted(set(owners.tolist())) == list(range(world))
    pick = [5, 1500, 2999]
    D, I = st.search(x[pick], 1)
    assert I[:, 0].tolist() == labels[pick].tolist() and np.allclose(D[:, 0], 1.0, atol=1e-5)
    st.remove(labels[pick[:2]])
    D, I = st.search(x[pick], 4)
    assert not (set(I.ravel().tolist()) & set(labels[pick[:2]].tolist()))
    assert I[2, 0] == labels[2999]
    more = st.add_balanced(x[:10] * np.float32(1.0), normalize=False)
    assert len(set(more.tolist()) | set(labels.tolist())) == 3010   # fresh labels, nothing renumbered
    st.close()


def _run_dead_peer(rank, world):
    """A rank that never issues the search: the others give up after the exchange timeout and raise
    (ExchangeTimeout) instead of spinning; the whole thing takes well under a second."""
    import torch.distributed as dist
    from minivectordb_b200.distributed import ExchangeTimeout, RowShardedIndex
    from oracle import oracle as O
    d = 64
    x = O.synth_rows(61, 0, 2000, d)
    idx = RowShardedIndex(d, device=rank, exchange="fused")
    idx.add(x[rank::world], normalize=True)
    idx.search(x[:1], 5)                     # healthy round first
    idx.set_exchange_timeout(250)
    dist.barrier()
    if rank == world - 1:
        raised = True                        # the "dead" rank: skips the round
    else:
        t0 = time.perf_counter()
        try:
            idx.search(x[:1], 5)
            raised = False
        except ExchangeTimeout:
            raised = True
        assert time.perf_counter() - t0 < 1.0
    assert raised
    idx.close()


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        for transport in (("fused", "nccl") if world > 1 else ("auto",)):
            _run_transport(rank, world, transport)
        _run_stable_numbering(rank, world)
        _run_shadow_mode(rank, world)
        if world > 1:
            _run_dead_peer(rank, world)
        ret[rank] = "ok"
    finally:
        dist.destroy_process_group()


def test_row_sharded_nccl():
    from minivectordb_b200 import _native
    world = min(8, _native.device_count())
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29700 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert dict(ret) == {r: "ok" for r in range(world)}
