"""CPU, world_size 2, gloo: the exchange logic of the row-sharded search
(offsets, label globalisation, mask slicing, all-gather layout, merge order)
on oracle-backed engines.  The CUDA scan and merge kernel are covered by the
-m gpu tests; here only the host plumbing of the N>1 path runs."""
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def merge_numpy(D_parts, I_parts, k):
    """Reference merge: score descending, then (part, position) ascending."""
    world, nq, kk = D_parts.shape
    D = np.full((nq, k), np.finfo(np.float32).min, dtype=np.float32)
    I = np.full((nq, k), -1, dtype=np.int64)
    for q in range(nq):
        cand = [(-float(D_parts[p, q, j]), p, j) for p in range(world) for j in range(kk) if I_parts[p, q, j] >= 0]
        cand.sort()
        for out, (_, p, j) in enumerate(cand[:k]):
            D[q, out] = D_parts[p, q, j]
            I[q, out] = I_parts[p, q, j]
    return D, I


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from fake_engine import FakeEngine
        from minivectordb_b200.distributed import RowShardedIndex
        from oracle import oracle as O
        n, d, k = 5000, 48, 10
        x = O.synth_rows(11, 0, n, d)
        O.normalize_L2(x)
        q = O.synth_rows(12, 0, 3, d)
        O.normalize_L2(q)
        bounds = [0, 1800, n]  # uneven shards
        idx = RowShardedIndex(d, engine_factory=FakeEngine, host_merge=merge_numpy)
        idx.add(x[bounds[rank]:bounds[rank + 1]], normalize=False)
        assert idx.offset == bounds[rank] and idx.ntotal_global == n
        D, I = idx.search(q, k)
        Dr, Ir = O.search_flat_ip(x, q, k)
        assert np.array_equal(I, Ir) and np.allclose(D, Dr, rtol=1e-6)
        adm = np.random.default_rng(0).random(n) < 0.2
        D, I = idx.search(q, k, mask_local=adm[bounds[rank]:bounds[rank + 1]])
        Dr, Ir = O.search_masked(x, adm, q, k)
        assert np.array_equal(I, Ir)
        # packed-filter entry point (what bench.py's multi-GPU e2e leg calls): this rank's slice as 32-bit words
        loc = adm[bounds[rank]:bounds[rank + 1]]
        words = np.zeros((loc.size + 31) // 32 * 4, dtype=np.uint8)
        pk = np.packbits(loc, bitorder="little")
        words[:pk.size] = pk
        Dp, Ip = idx.search_packed(q, k, words, loc.size)
        assert np.array_equal(Ip, Ir) and np.array_equal(Dp, D)
        Du, Iu = idx.search_packed(q[:1], k)
        assert np.array_equal(Iu, O.search_flat_ip(x, q[:1], k)[1])
        # k larger than one shard's admissible rows: padding must not leak into the merge
        few = np.zeros(n, dtype=bool)
        few[[5, 1700, 1801, 4999]] = True
        D, I = idx.search(q, k, mask_local=few[bounds[rank]:bounds[rank + 1]])
        assert sorted(I[0][I[0] >= 0].tolist()) == [5, 1700, 1801, 4999] and (I[0][4:] == -1).all()
        # a second collective add keeps global numbering contiguous per rank... and is visible
        idx.add(x[:10] if rank == 1 else None, normalize=False)
        assert idx.ntotal_global == n + 10
        # stable numbering: balanced inserts (rows go to the least-full rank) and deletes by label
        st = RowShardedIndex(d, engine_factory=FakeEngine, host_merge=merge_numpy, numbering="stable")
        st.add(x[:7] if rank == 0 else None, normalize=False)        # rank 0 starts 7 rows ahead
        lab0 = np.arange(7, dtype=np.int64)
        labels = st.add_balanced(x[7:107], normalize=False)           # same block on every rank
        owners = labels >> RowShardedIndex.STABLE_SHIFT
        assert (owners[:7] == 1).all()                                # rank 1 catches up first ...
        assert abs(int((owners == 0).sum()) + 7 - int((owners == 1).sum())) <= 1   # ... then they alternate
        assert st.counts == [7 + int((owners == 0).sum()), int((owners == 1).sum())]
        all_labels = np.concatenate([lab0, labels])
        D, I = st.search(x[:107][[3, 50, 99]], 1)
        assert I[:, 0].tolist() == all_labels[[3, 50, 99]].tolist() and np.allclose(D[:, 0], 1.0, atol=1e-5)
        st.remove(all_labels[[50, 99]])
        D, I = st.search(x[:107][[50, 99]], 3)
        assert not (set(I.ravel().tolist()) & set(all_labels[[50, 99]].tolist()))
        with pytest.raises(ValueError):
            idx.add_balanced(x[:2])                                   # contiguous numbering cannot do it
        ret[rank] = "ok"
    finally:
        dist.destroy_process_group()


def test_row_sharded_exchange_world2():
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert dict(ret) == {0: "ok", 1: "ok"}
