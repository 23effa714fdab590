"""GPU: the drop-in classes on the real CUDA engine -- the golden scenario the
REFERENCE's classes produced, and mirrors of the reference's own tests."""
import numpy as np
import pytest

import dropin_cases as C

pytestmark = pytest.mark.gpu


@pytest.fixture()
def classes():
    from minivectordb_b200 import VectorDatabase, ShardedVectorDatabase
    return VectorDatabase, ShardedVectorDatabase


def _ndev():
    from minivectordb_b200 import _native
    return _native.device_count()


def test_golden_scenario_vdb(classes, tmp_path):
    C.case_golden_scenario_vdb(classes[0], tmp_path)


def test_golden_scenario_svdb(classes, tmp_path):
    C.case_golden_scenario_svdb(classes[1], tmp_path)


def _devices(n):
    """n partitions over the visible GPUs (several partitions share a GPU on a small box: the fused
    exchange still runs between them, through the same peer-store code path)."""
    return [i % _ndev() for i in range(n)]


@pytest.mark.parametrize("nparts", [2, 3, 8])
def test_golden_scenario_svdb_shard_group(classes, tmp_path, nparts):
    """ShardedVectorDatabase(devices=[...]): ONE fused group search per query (scan + NVLink exchange +
    merge inside the kernels), same answers as the reference's single in-memory index."""
    import minivectordb_b200._store as store
    calls = []
    real = store.FlatIPEngine.group_class.search

    def spy(self, *a, **kw):
        calls.append(1)
        return real(self, *a, **kw)

    store.FlatIPEngine.group_class.search = spy
    try:
        C.case_golden_scenario_svdb(classes[1], tmp_path, devices=_devices(nparts))
    finally:
        store.FlatIPEngine.group_class.search = real
    assert calls, "the multi-device class did not use the shard group"


def test_sharded_basics_shard_group(classes, tmp_path):
    C.case_sharded_basics(classes[1], tmp_path, devices=_devices(2))


def test_golden_scenario_svdb_partitions_in_parallel(classes, tmp_path, monkeypatch):
    # fallback without a shard group (also what k > 128 takes): partitions scanned from worker threads
    import minivectordb_b200._store as store
    monkeypatch.setattr(store.GpuStore, "PARALLEL_PARTS_BYTES", 0)
    monkeypatch.setattr(store.FlatIPEngine, "group_class", None)
    C.case_golden_scenario_svdb(classes[1], tmp_path, devices=_devices(2))


def test_loads_reference_pickle(classes, tmp_path):
    C.case_loads_reference_pickle(classes[0], tmp_path)


def test_basics(classes, tmp_path):
    C.case_basics(classes[0], tmp_path)


def test_exclude_enumerates_duplicates(classes, tmp_path):
    C.case_exclude_enumerates_duplicates(classes[0], tmp_path)


def test_mongolike(classes, tmp_path):
    C.case_mongolike(classes[0], tmp_path)
    C.case_mongolike(lambda storage_file: classes[1](storage_dir=storage_file + "_d"), tmp_path)


def test_autocut_and_rerank(classes, tmp_path):
    C.case_autocut_and_rerank(classes[0], tmp_path)


def test_sharded_basics(classes, tmp_path):
    C.case_sharded_basics(classes[1], tmp_path)


def test_migration(classes, tmp_path):
    C.case_migration(classes[0], classes[1], tmp_path)


def test_multithreaded_full_size(classes, tmp_path):
    # ref tests/test_multithreaded_operations.py at its own sizes (5000 + 5x2000 inserts, 3500 deletes)
    C.case_multithreaded(classes[0], tmp_path, scale=1.0)


def test_sharded_multithreaded_with_large_k(classes, tmp_path):
    """ref tests/test_sharded_multithreaded_operations.py:12-107 (d=512, k=825 search under churn), reduced rows."""
    import threading
    import uuid
    db = classes[1](storage_dir=str(tmp_path / "mt"), shard_size=777)
    d = 512
    ids = [str(uuid.uuid4()) for _ in range(3000)]
    db.store_embeddings_batch(ids, [np.random.rand(d) for _ in ids], [{"f": i} for i in range(3000)])
    errors = []

    def ins():
        try:
            for _ in range(200):
                db.store_embedding(str(uuid.uuid4()), np.random.rand(d), {"f": -1})
        except Exception as e:  # noqa: BLE001
            errors.append(e)

    def srch():
        try:
            for i in range(100):
                out = db.find_most_similar(np.random.rand(d), k=825 if i % 10 == 0 else 5)
                assert len(out[0]) == len(out[1]) == len(out[2])
        except Exception as e:  # noqa: BLE001
            errors.append(e)

    def dele():
        try:
            for a in range(100, 300, 20):
                db.delete_embeddings_batch(ids[a:a + 20])
        except Exception as e:  # noqa: BLE001
            errors.append(e)

    ts = [threading.Thread(target=f) for f in (ins, ins, srch, srch, srch, dele)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errors, errors[:1]
    assert len(db.unique_ids) == len(db.inverse_id_map) == len(db.metadata) == len(db.embeddings) == 3000 + 400 - 200


def test_compaction_on_device(classes, tmp_path, monkeypatch):
    import minivectordb_b200._store as store
    monkeypatch.setattr(store.GpuStore, "COMPACT_MIN_DEAD", 8)
    db = classes[0](storage_file=str(tmp_path / "cmp.pkl"))
    rng = np.random.default_rng(0)
    embs = rng.standard_normal((2000, 48)).astype(np.float32)
    db.store_embeddings_batch(list(range(2000)), list(embs), [{"g": i % 4} for i in range(2000)])
    before = db.find_most_similar(embs[777], k=20, metadata_filter={"g": {"$ne": 0}})
    for i in range(0, 2000, 4):
        db.delete_embedding(i)   # g == 0 rows
    after = db.find_most_similar(embs[777], k=20, metadata_filter={"g": {"$ne": 0}})
    assert db._parts[0].engine.ntotal == 1500   # physically compacted
    assert before[0] == after[0] and np.array_equal(np.array(before[1]), np.array(after[1]))


def test_device_filter_evaluation_matches_host_evaluation(classes, tmp_path):
    """Numeric predicates are evaluated by CUDA kernels over HBM-resident metadata columns and
    combined on the device; the admissible set must equal the host (numpy) evaluation for every
    combinator, including keys that are missing, mixed-type columns and rows appended later."""
    from minivectordb_b200.filters import FilterIndex
    db = classes[0](storage_file=str(tmp_path / "dev.pkl"))
    rng = np.random.default_rng(2)
    n, d = 5000, 16
    embs = rng.standard_normal((n, d)).astype(np.float32)
    metas = []
    for i in range(n):
        m = {"tag": f"t{i % 7}"}
        if i % 3:
            m["value"] = int(rng.integers(0, 100))
        if i % 5 == 0:
            m["score"] = float(rng.random())
        if i % 11 == 0:
            m["mixed"] = i if i % 2 else str(i)
        metas.append(m)
    db.store_embeddings_batch(list(range(n)), list(embs), metas)
    q = rng.standard_normal(d).astype(np.float32)
    filters = [
        dict(metadata_filter={"value": {"$gt": 49}}),
        dict(metadata_filter={"value": {"$lte": 10}, "score": {"$lt": 0.5}}),
        dict(metadata_filter={"value": {"$ne": 7}}),
        dict(metadata_filter={"value": 42}),
        dict(metadata_filter=[{"value": {"$gte": 20}}, {"tag": "t3"}]),
        dict(or_filters=[{"value": {"$lt": 5}}, {"score": {"$gt": 0.9}}, {"nokey": 1}]),
        dict(metadata_filter={"value": {"$gt": 10}}, or_filters={"tag": "t1", "score": {"$gt": 0.5}}),
        dict(exclude_filter={"value": 42}),
        dict(metadata_filter={"value": {"$gt": 30}}, exclude_filter=[{"tag": "t2"}, {"value": 77}]),
        dict(metadata_filter={"tag": {"$ne": "t3"}}),                      # string columns are dictionary-coded on the device
        dict(metadata_filter={"tag": "never-stored"}),
        dict(metadata_filter={"value": {"$lt": 90}}, exclude_filter={"tag": "never-stored"}),
        dict(metadata_filter={"tag": {"$ne": "never-stored"}, "value": {"$gt": 95}}),
        dict(metadata_filter={"mixed": {"$ne": 11}}),
        dict(metadata_filter={"nokey": {"$gt": 1}}),
    ]
    db.find_most_similar(q, k=1)   # flush
    db.delete_embedding(5)
    db.delete_embedding(4242)

    def check_all():
        for f in filters:
            got = db.find_most_similar(q, k=10_000, **f)
            live = db._g_live[:db._g_n]
            want = FilterIndex.admissible(db._filters, live, f.get("metadata_filter"), f.get("exclude_filter"),
                                          f.get("or_filters"))
            want_ids = set(np.flatnonzero(live if want is None else want).tolist())
            assert set(db.inverse_id_map[u] for u in got[0]) == {int(np.sum(live[:g])) for g in want_ids}, f
            # a repeat is served from the cached device-resident mask
            again = db.find_most_similar(q, k=10_000, **f)
            assert again[0] == got[0]

    check_all()
    # rows appended after the device columns were built extend them lazily
    more = rng.standard_normal((300, d)).astype(np.float32)
    db.store_embeddings_batch(list(range(n, n + 300)), list(more), [{"value": int(i % 100), "tag": "t3"} for i in range(300)])
    check_all()
    with pytest.raises(ValueError):
        db.find_most_similar(q, metadata_filter={"value": {"$regex": 1}})


def test_golden_scenarios_with_the_int8_shadow_mode(classes, tmp_path):
    """scan_shadow=True is an exact mode: the golden scenarios (recorded from the reference's own classes) still
    hold, on one device and on a shard group (rows below the mode's 16384-row floor simply take the fp32 scan)."""
    C.case_golden_scenario_vdb(lambda **kw: classes[0](scan_shadow=True, **kw), tmp_path)
    C.case_golden_scenario_svdb(lambda **kw: classes[1](scan_shadow=True, **kw), tmp_path, devices=_devices(2))
    # a database large enough for the shadow pass to engage
    import numpy as np
    rng = np.random.default_rng(0)
    emb = rng.standard_normal((40_000, 96)).astype(np.float32)
    a = classes[0](storage_file=str(tmp_path / "a.pkl"))
    b = classes[0](storage_file=str(tmp_path / "b.pkl"), scan_shadow=True)
    for db in (a, b):
        db.store_embeddings_batch(list(range(40_000)), emb, [{"v": i % 7} for i in range(40_000)])
    for i in (0, 17, 39_999):
        ra = a.find_most_similar(emb[i] + 0.01, k=10)
        rb = b.find_most_similar(emb[i] + 0.01, k=10)
        assert ra[0] == rb[0] and np.array_equal(np.asarray(ra[1]), np.asarray(rb[1]))
        ra = a.find_most_similar(emb[i], k=5, metadata_filter={"v": {"$gt": 3}})
        rb = b.find_most_similar(emb[i], k=5, metadata_filter={"v": {"$gt": 3}})
        assert ra[0] == rb[0] and np.array_equal(np.asarray(ra[1]), np.asarray(rb[1]))
    a.close(); b.close()
