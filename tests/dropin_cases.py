"""Behaviour checks of the drop-in classes, written once and run twice: on the
oracle-backed FakeEngine (CPU, host logic) and on the real CUDA engine (GPU).
Each case cites the reference test it mirrors (ref: /root/reference/tests/)."""
import json
import os
import sys
import threading
import uuid
from datetime import datetime

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import scenario  # noqa: E402


def _golden(name):
    return json.load(open(os.path.join(HERE, "golden", name)))


def _compare_queries(got, want):
    assert len(got) == len(want)
    for i, (g, w) in enumerate(zip(got, want)):
        assert g["ids"] == w["ids"], (i, g["ids"][:5], w["ids"][:5])
        assert g["n_meta"] == w["n_meta"]
        assert np.allclose(g["dist"], w["dist"], rtol=1e-5, atol=1e-6), i


def case_golden_scenario_vdb(VDB, tmp_path):
    """Same scenario the REFERENCE VectorDatabase produced tests/golden/reference_vdb_scenario.json from."""
    rec = scenario.run(VDB, storage_file=str(tmp_path / "none.pkl"))
    want = _golden("reference_vdb_scenario.json")
    _compare_queries(rec["queries_before"], want["queries_before"])
    _compare_queries(rec["queries_after"], want["queries_after"])
    assert [list(x) for x in rec["id_map"]] == want["id_map"]
    assert [list(x) for x in rec["inverse_id_map"]] == want["inverse_id_map"]
    assert rec["n_metadata"] == want["n_metadata"] and rec["n_embeddings"] == want["n_embeddings"]
    assert np.allclose(rec["raw_vector_before_search"], want["raw_vector_before_search"])


def case_golden_scenario_svdb(SVDB, tmp_path, **kw):
    rec = scenario.run(SVDB, sharded=True, storage_dir=str(tmp_path / "shards"), shard_size=64, **kw)
    want = _golden("reference_svdb_scenario.json")
    _compare_queries(rec["queries_before"], want["queries_before"])
    _compare_queries(rec["queries_after"], want["queries_after"])
    assert [list(x) for x in rec["inverse_id_map"]] == want["inverse_id_map"]
    assert rec["n_metadata"] == want["n_metadata"] and rec["n_embeddings"] == want["n_embeddings"]


def case_loads_reference_pickle(VDB, tmp_path):
    """A db.pkl written by the REFERENCE loads here (ref vector_database.py:28-40, 538-548)."""
    import shutil
    path = str(tmp_path / "ref.pkl")
    shutil.copy(os.path.join(HERE, "golden", "reference_db.pkl"), path)
    db = VDB(storage_file=path)
    emb, meta, ids, _, queries = scenario.make_rows()
    assert len(db.id_map) == 40 and db.embedding_size == scenario.D
    assert db.id_map[3] == ids[3] and db.metadata[3] == meta[3]
    out = db.find_most_similar(emb[7], k=1)
    assert out[0][0] == ids[7] and abs(out[1][0] - 1.0) < 1e-5


def case_basics(VDB, tmp_path):
    f = str(tmp_path / "a.pkl")
    db = VDB(storage_file=f)
    # fresh state (ref test_vector_database.py:7-11)
    assert db.embedding_size is None and len(db.id_map) == 0 and len(db.inverse_id_map) == 0
    # empty DB search (ref :162-168)
    assert db.find_most_similar([1.0, 0.0], k=3) == ([], [], [])
    # first insert fixes the size, d=2 works (ref :13-18)
    db.store_embedding(1, [0.5, 0.5])
    assert db.embedding_size == 2
    # get_vector returns the stored values before any search (ref :365-372)
    assert (db.get_vector(1) == [0.5, 0.5]).all()
    db.store_embedding(2, [0.1, 0.1])
    db.store_embedding(3, [0.2, 0.2])
    # duplicate / unknown ids (ref :325-347, 374-380)
    with pytest.raises(ValueError):
        db.store_embedding(1, [0.3, 0.3])
    with pytest.raises(ValueError):
        db.delete_embedding(99)
    with pytest.raises(ValueError):
        db.get_vector(99)
    with pytest.raises(ValueError):
        db.store_embeddings_batch([7, 1], [[0.1, 0.2], [0.3, 0.4]])
    with pytest.raises(ValueError):
        db.store_embeddings_batch([7, 8], [[0.1, 0.2], [0.3, 0.4]], [{"a": 1}])
    with pytest.raises(ValueError):
        db.store_embedding(50, [0.1, 0.2, 0.3])  # wrong dimension
    # a batch with one ragged / wrong-sized row stores NOTHING (block conversion happens first)
    before = (dict(db.id_map), len(db.metadata))
    with pytest.raises(ValueError):
        db.store_embeddings_batch([60, 61], [[0.1, 0.2], [0.3, 0.4, 0.5]])
    with pytest.raises(ValueError):
        db.store_embeddings_batch([60, 61], np.zeros((2, 3), dtype=np.float32))
    assert (dict(db.id_map), len(db.metadata)) == before
    # batches as a list of rows and as one 2-D array land identically (bulk append path)
    db.store_embeddings_batch([70, 71], [np.array([0.6, 0.1]), [0.2, 0.7]], [{"b": 1}, {"b": 2}])
    db.store_embeddings_batch([72, 73], np.array([[0.3, 0.3], [0.9, 0.2]], dtype=np.float64))
    assert (db.get_vector(71) == np.array([0.2, 0.7], dtype=np.float32)).all()
    assert (db.get_vector(73) == np.array([0.9, 0.2], dtype=np.float32)).all()
    assert db.metadata[db.inverse_id_map[70]] == {"b": 1} and db.metadata[db.inverse_id_map[72]] == {}
    assert db.inverted_index["b"] == {70, 71}
    assert db.find_most_similar([0.2, 0.7], metadata_filter={"b": {"$gt": 1}}, k=3)[0] == (71,)
    for uid in (70, 71, 72, 73):
        db.delete_embedding(uid)
    # delete renumbers densely (ref :349-363)
    db.delete_embedding(2)
    assert db.id_map == {0: 1, 1: 3} and db.inverse_id_map == {1: 0, 3: 1}
    # k > N returns N, tuples, np.float32 scores descending (ref :149-160)
    ids, dist, meta = db.find_most_similar([1.0, 0.5], k=10)
    assert isinstance(ids, tuple) and len(ids) == 2 and isinstance(dist[0], np.float32)
    assert dist[0] >= dist[1]
    # after a search the stored rows are normalised (ref vector_database.py:45)
    assert abs(np.linalg.norm(db.get_vector(1)) - 1.0) < 1e-6
    # persist + reload (ref :177-193)
    db.store_embedding("x", [0.9, -0.1], {"kind": "neg"})
    db.persist_to_disk()
    db2 = VDB(storage_file=f)
    assert db2.id_map == db.id_map and db2.metadata == db.metadata
    assert db2.find_most_similar([0.9, -0.1], k=1)[0] == ("x",)
    assert len(db2.embeddings) == 3 and db2.embeddings.shape == (3, 2)


def case_exclude_enumerates_duplicates(VDB, tmp_path):
    """Repeated k=1 with a growing exclude list visits every row once, even with
    duplicate vectors (ref test_vector_database.py:34-97)."""
    db = VDB(storage_file=str(tmp_path / "b.pkl"))
    rng = np.random.default_rng(0)
    base = rng.random(8)
    for i in range(12):
        db.store_embedding(i, base if i < 6 else rng.random(8), {"uid": i})
    seen, excl = [], []
    for _ in range(12):
        ids, _, _ = db.find_most_similar(base, exclude_filter=excl or None, k=1)
        seen.append(ids[0])
        excl.append({"uid": ids[0]})
    assert sorted(seen) == list(range(12))
    assert db.find_most_similar(base, exclude_filter=excl, k=1) == ([], [], [])


def case_mongolike(VDB, tmp_path):
    """Operator semantics incl. datetime, $in, invalid operator, contradiction
    (ref test_mongolike_operators.py:9-248)."""
    db = VDB(storage_file=str(tmp_path / "c.pkl"))
    rng = np.random.default_rng(1)
    for i in range(250):
        db.store_embedding(f"item_{i}", rng.random(4), {"num_filter": f"test_{int(rng.integers(1, 5))}"})
    for i in range(10):
        db.store_embedding(f"item_{250 + i}", rng.random(4),
                           {"num_filter": "test_10", "value": 10, "date": datetime(2021, 1, 1)})
    for i in range(10):
        db.store_embedding(f"item_{260 + i}", rng.random(4),
                           {"num_filter": "test_20", "value": 20, "date": datetime(2022, 1, 1), "ids": [i, 100 + i]})
    q = rng.random(4)

    def n(**kw):
        return len(db.find_most_similar(q, k=999, **kw)[0])

    assert n(metadata_filter={"value": 10}) == 10
    assert n(metadata_filter={"value": {"$gte": 10}}) == 20
    assert n(metadata_filter={"value": {"$gte": 20}}) == 10
    assert n(metadata_filter={"value": {"$lt": 20}}) == 10
    assert n(metadata_filter={"value": {"$lte": 10}}) == 10
    assert n(metadata_filter={"value": {"$ne": 10}}) == 10        # rows lacking the key do not match
    assert n(metadata_filter={"date": {"$gte": datetime(2021, 1, 1)}}) == 20
    assert n(metadata_filter={"date": {"$lt": datetime(2022, 1, 1)}}) == 10
    assert n(metadata_filter={"value": {"$gt": 15}, "date": {"$gt": datetime(2021, 5, 5)}}) == 10
    assert n(or_filters=[{"value": {"$gte": 10}}, {"date": {"$lte": datetime(2022, 1, 1)}}]) == 20
    assert n(metadata_filter=[{"value": {"$gte": 10}}, {"date": {"$lte": datetime(2021, 6, 1)}}]) == 10
    assert n(metadata_filter={"ids": {"$in": 3}}) == 1             # operand contained in the stored list
    assert n(metadata_filter={"ids": {"$in": 103}}) == 1
    assert n(metadata_filter={"value": {"$gt": 15}}, or_filters={"value": {"$lt": 15}}) == 0
    assert n(metadata_filter={"value": {"$gte": 0, "$lte": 5}}) == 20  # only the first operator is honoured
    with pytest.raises(ValueError):
        db.find_most_similar(q, metadata_filter={"value": {"$regex": 1}})
    with pytest.raises(ValueError):
        db.find_most_similar(q, or_filters={"value": {"$nope": 1}})
    res = db.find_most_similar(q, k=999, metadata_filter={"value": {"$gte": 10}})
    assert all(m["value"] >= 10 for m in res[2])


def case_autocut_and_rerank(VDB, tmp_path):
    db = VDB(storage_file=str(tmp_path / "d.pkl"))
    assert db.autocut_scores([0.9, 0.88, 0.5, 0.49]) == [2, 3]
    assert db.autocut_scores([0.9, 0.85, 0.8]) == []
    # autocut trims and turns the containers into lists (ref vector_database.py:528-534)
    db.store_embedding("a", [1.0, 0.0])
    db.store_embedding("b", [0.99, 0.05])
    db.store_embedding("c", [0.0, 1.0])
    ids, dist, meta = db.find_most_similar([1.0, 0.0], k=3, autocut=True)
    assert ids == ["a", "b"] and isinstance(dist, list)
    # empty rerank input (ref test_vector_database.py:554-570)
    s, sc = db.hybrid_rerank_results([], [], "query", k=5)
    assert len(s) == 0 and len(sc) == 0
    sents = ["the cat sat on the mat", "quantum chromodynamics lecture", "a cat on a mat"]
    s, sc = db.hybrid_rerank_results(sents, [0.5, 0.5, 0.5], "cat mat", k=2)
    assert len(s) == 2 and "quantum chromodynamics lecture" not in s


def case_sharded_basics(SVDB, tmp_path, **kw):
    d = str(tmp_path / "s1")
    db = SVDB(storage_dir=d, shard_size=50, **kw)
    assert db.embedding_size is None and len(db.inverse_id_map) == 0
    assert db.find_most_similar([0.1, 0.2], k=2) == ([], [], [])
    rng = np.random.default_rng(3)
    ids = list(range(230))
    embs = [rng.random(16) for _ in ids]
    db.store_embeddings_batch(ids, embs, [{"v": i % 10} for i in ids])
    with pytest.raises(ValueError):
        db.store_embeddings_batch([1000, 1001], [embs[0]])
    with pytest.raises(ValueError):
        db.store_embeddings_batch([5], [embs[0]])
    assert len([f for f in os.listdir(d) if f.endswith(".pkl")]) == 5
    # deletes: scalar, list, errors (ref test_sharded_vector_database.py:369-384, 613-642)
    db.delete_embeddings_batch(2)
    db.delete_embeddings_batch([10, 60, 120])
    for bad in ([], None, [3, 99999], [None]):
        with pytest.raises(ValueError):
            db.delete_embeddings_batch(bad)
    assert len(db.inverse_id_map) == 226 and db.inverse_id_map[0] == 0 and db.inverse_id_map[3] == 2
    assert db.unique_ids[:4] == [0, 1, 3, 4]
    # k=500 through a filter returns exactly the admissible count (ref :663-694)
    got = db.find_most_similar(embs[7], k=500, metadata_filter={"v": 7})
    assert len(got[0]) == 23 and got[0][0] == 7
    # raw vector comes back from the shard file
    assert np.allclose(db.get_vector(7), embs[7].astype(np.float32))
    # reload from the shard files
    db2 = SVDB(storage_dir=d, shard_size=50, **kw)
    assert db2.unique_ids == db.unique_ids and db2.metadata == db.metadata
    a = db.find_most_similar(embs[33], k=5)
    b = db2.find_most_similar(embs[33], k=5)
    assert a[0] == b[0] and np.allclose(a[1], b[1], atol=1e-6)
    # delete half, reload, everything still consistent; then delete all
    db2.delete_embeddings_batch([i for i in db2.unique_ids if i % 2 == 0])
    db3 = SVDB(storage_dir=d, shard_size=50, **kw)
    assert db3.unique_ids == db2.unique_ids and len(db3.find_most_similar(embs[1], k=500)[0]) == len(db3.unique_ids)
    db3.delete_embeddings_batch(list(db3.unique_ids))
    assert db3.find_most_similar(embs[1], k=5) == ([], [], [])
    assert len(SVDB(storage_dir=d, shard_size=50, **kw).unique_ids) == 0


def case_migration(VDB, SVDB, tmp_path, **kw):
    """_convert_from_non_sharded_db (ref test_sharded_vector_database.py:644-661)."""
    v = VDB(storage_file=str(tmp_path / "m.pkl"))
    rng = np.random.default_rng(5)
    for i in range(40):
        v.store_embedding(f"id{i}", rng.random(8), {"i": i})
    s = SVDB(storage_dir=str(tmp_path / "ms"), shard_size=16, **kw)
    s._convert_from_non_sharded_db(v)
    assert s.unique_ids == [f"id{i}" for i in range(40)]
    q = rng.random(8)
    a, b = v.find_most_similar(q, k=7), s.find_most_similar(q, k=7)
    assert a[0] == b[0] and np.allclose(a[1], b[1], atol=1e-6)


def case_multithreaded(VDB, tmp_path, scale=1.0):
    """Concurrent store / search / delete ends with consistent sizes
    (ref test_multithreaded_operations.py:4-62)."""
    db = VDB(storage_file=str(tmp_path / "t.pkl"))
    d, initial = 64, int(5000 * scale)
    n_ins, n_search = int(2000 * scale), int(1000 * scale)
    del_lo, del_hi = int(500 * scale), int(4000 * scale)
    db.store_embeddings_batch(list(range(initial)), [np.random.rand(d) for _ in range(initial)],
                              [{"num_filter": f"test_{i}"} for i in range(initial)])
    errors = []

    def guard(fn):
        def run():
            try:
                fn()
            except Exception as e:  # noqa: BLE001
                errors.append(e)
        return run

    def index_thread():
        for i in range(n_ins):
            db.store_embedding(f"item_{uuid.uuid4()}", np.random.rand(d), metadata_dict={"num_filter": f"test_{i}"})

    def search_thread():
        for _ in range(n_search):
            ids, dist, meta = db.find_most_similar(embedding=np.random.rand(d), k=3)
            assert len(ids) == len(dist) == len(meta) <= 3

    def delete_thread():
        for i in range(del_lo, del_hi):
            db.delete_embedding(i)

    threads = [threading.Thread(target=guard(f)) for _ in range(5) for f in (index_thread, search_thread)]
    threads.append(threading.Thread(target=guard(delete_thread)))
    [t.start() for t in threads]
    [t.join() for t in threads]
    assert not errors, errors[:1]
    want = initial + 5 * n_ins - (del_hi - del_lo)
    assert len(db.id_map) == len(db.inverse_id_map) == len(db.metadata) == len(db.embeddings) == want
    ids, dist, _ = db.find_most_similar(np.random.rand(d), k=5)
    assert len(ids) == 5 and all(dist[i] >= dist[i + 1] for i in range(4))
