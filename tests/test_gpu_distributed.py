"""GPU: row-sharded search, one process per GPU over NCCL (world = min(2, visible GPUs)).
With a single visible GPU the same code path runs at world 1."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _run_transport(rank, world, transport):
    from minivectordb_b200.distributed import RowShardedIndex
    from oracle import oracle as O
    n, d, k = 40000, 384, 10
    x = O.synth_rows(11, 0, n, d)
    O.normalize_L2(x)
    q = O.synth_rows(12, 0, 4, d)
    O.normalize_L2(q)
    bounds = np.linspace(0, n, world + 1).astype(int)
    bounds[1:-1] += 37  # uneven shards
    idx = RowShardedIndex(d, device=rank, exchange=transport)
    assert world == 1 or idx.exchange == transport
    idx.add(x[bounds[rank]:bounds[rank + 1]], normalize=False)
    assert idx.offset == bounds[rank] and idx.ntotal_global == n
    for kk in (1, k, 100):
        D, I = idx.search(q, kk)
        Dr, Ir = O.search_flat_ip(x, q, kk)
        rep = O.classify_parity(x, q, I, D, Ir, Dr)
        assert rep["ok"], (transport, kk, rep)
    adm = np.random.default_rng(0).random(n) < 0.3
    for _ in range(20):  # many rounds: sequence numbers / parity double-buffering
        D, I = idx.search(q, k, mask_local=adm[bounds[rank]:bounds[rank + 1]])
    Dr, Ir = O.search_masked(x, adm, q, k)
    rep = O.classify_parity(x, q, I, D, Ir, Dr, admissible=adm)
    assert rep["ok"], (transport, rep)
    # a shard with nothing admissible still takes part; 11 queries = groups of 8 + 2 + 1
    q11 = O.synth_rows(13, 0, 11, d)
    O.normalize_L2(q11)
    only0 = np.zeros(n, dtype=bool)
    only0[:50] = True
    D, I = idx.search(q11, k, mask_local=only0[bounds[rank]:bounds[rank + 1]])
    Dr, Ir = O.search_masked(x, only0, q11, k)
    rep = O.classify_parity(x, q11, I, D, Ir, Dr, admissible=only0)
    assert rep["ok"], (transport, rep)
    # host-buffer fast path (one H2D, one D2H) and pipelined launches (option "pdl") give the same answers
    import minivectordb_b200 as mv
    loc = adm[bounds[rank]:bounds[rank + 1]]
    words = np.zeros((loc.size + 31) // 32 * 4, dtype=np.uint8)
    pk = mv.pack_mask(loc)
    words[:pk.size] = pk
    Dm, Im = idx.search(q, k, mask_local=loc)
    for pdl in (0, 1):
        idx.engine.set_option("pdl", pdl)
        for i in range(12):
            Dp, Ip = idx.search_packed(q[i % 4:i % 4 + 1], k, words, loc.size)
            assert np.array_equal(Ip[0], Im[i % 4]) and np.array_equal(Dp[0], Dm[i % 4]), (transport, pdl, i)
        Du, Iu = idx.search_packed(q, kk)
        Dr, Ir = O.search_flat_ip(x, q, kk)
        assert O.classify_parity(x, q, Iu, Du, Ir, Dr)["ok"]
    idx.engine.set_option("pdl", 0)
    assert not idx.exchange_timed_out()
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
        ret[rank] = "ok"
    finally:
        dist.destroy_process_group()


def test_row_sharded_nccl():
    from minivectordb_b200 import _native
    world = min(2, _native.device_count())
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29700 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert dict(ret) == {r: "ok" for r in range(world)}
