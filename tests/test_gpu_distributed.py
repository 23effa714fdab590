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


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from minivectordb_b200.distributed import RowShardedIndex
        from oracle import oracle as O
        n, d, k = 40000, 384, 10
        x = O.synth_rows(11, 0, n, d)
        O.normalize_L2(x)
        q = O.synth_rows(12, 0, 4, d)
        O.normalize_L2(q)
        bounds = np.linspace(0, n, world + 1).astype(int)
        bounds[1:-1] += 37  # uneven shards
        idx = RowShardedIndex(d, device=rank)
        idx.add(x[bounds[rank]:bounds[rank + 1]], normalize=False)
        assert idx.offset == bounds[rank] and idx.ntotal_global == n
        for kk in (1, k, 100):
            D, I = idx.search(q, kk)
            Dr, Ir = O.search_flat_ip(x, q, kk)
            rep = O.classify_parity(x, q, I, D, Ir, Dr)
            assert rep["ok"], rep
        adm = np.random.default_rng(0).random(n) < 0.3
        D, I = idx.search(q, k, mask_local=adm[bounds[rank]:bounds[rank + 1]])
        Dr, Ir = O.search_masked(x, adm, q, k)
        rep = O.classify_parity(x, q, I, D, Ir, Dr, admissible=adm)
        assert rep["ok"], rep
        idx.close()
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
