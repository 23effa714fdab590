"""Deterministic scenario shared by make_golden.py (runs it on the REFERENCE's
classes) and the parity tests (run it on minivectordb_b200's classes)."""
from datetime import datetime

import numpy as np

D = 32
N_BATCH, N_SINGLE = 300, 100
TAGS = [f"s{i}" for i in range(16)]


def make_rows(seed=2024):
    rng = np.random.default_rng(seed)
    n = N_BATCH + N_SINGLE
    emb = rng.standard_normal((n, D)).astype(np.float32)
    meta = []
    for i in range(n):
        m = {"tag": f"t{int(rng.integers(0, 8))}",
             "tags": [TAGS[int(a)] for a in rng.choice(16, 2, replace=False)]}
        if i % 7 != 3:
            m["value"] = int(rng.integers(0, 100))
        if i % 5 == 0:
            m["date"] = datetime(2020 + int(rng.integers(0, 4)), 1 + int(rng.integers(0, 12)), 1)
        meta.append(m)
    ids = [f"item_{i}" if i % 2 else i for i in range(n)]  # mixed id types
    deletes = [ids[int(j)] for j in rng.choice(n, 37, replace=False)]
    queries = rng.standard_normal((24, D)).astype(np.float32)
    return emb, meta, ids, deletes, queries


def query_specs():
    """(kwargs) for find_most_similar; the query vector is queries[i]."""
    return [
        dict(k=10),
        dict(k=5),
        dict(k=1),
        dict(k=999),
        dict(k=10, metadata_filter={"value": {"$gt": 49}}),
        dict(k=10, metadata_filter={"value": {"$gte": 50}}),
        dict(k=10, metadata_filter={"value": {"$lt": 10}}),
        dict(k=999, metadata_filter={"value": {"$lte": 10}}),
        dict(k=10, metadata_filter={"value": {"$ne": 10}}),
        dict(k=10, metadata_filter={"value": 42}),
        dict(k=10, metadata_filter={"tags": {"$in": "s3"}}),
        dict(k=10, metadata_filter={"tag": "t3", "value": {"$gt": 20}}),
        dict(k=10, metadata_filter=[{"tag": "t1"}, {"value": {"$lt": 70}}]),
        dict(k=10, metadata_filter={"date": {"$gte": datetime(2022, 1, 1)}}),
        dict(k=10, or_filters=[{"tag": "t0"}, {"value": {"$gt": 90}}]),
        dict(k=10, or_filters={"tag": "t2", "value": 5}),
        dict(k=10, metadata_filter={"value": {"$gt": 30}}, or_filters=[{"tag": "t4"}, {"tag": "t5"}]),
        dict(k=10, exclude_filter={"tag": "t0"}),
        dict(k=10, exclude_filter=[{"tag": "t0"}, {"tag": "t1"}, {"value": 42}]),
        dict(k=10, metadata_filter={"value": {"$gt": 10}}, exclude_filter={"tag": "t7"}, or_filters={"tag": "t7"}),
        dict(k=10, metadata_filter={"tag": "nope"}),
        dict(k=500, metadata_filter={"tag": "t6"}),
        dict(k=10, metadata_filter={"value": {"$gte": 0, "$lte": 5}}),  # only the FIRST operator counts
        dict(k=10, metadata_filter={}, or_filters=[{}]),
    ]


def run(db_cls, sharded=False, **ctor):
    """Drive one database class through the scenario; returns a JSON-able record."""
    emb, meta, ids, deletes, queries = make_rows()
    db = db_cls(**ctor)
    rec = {"queries_before": [], "queries_after": []}
    db.store_embeddings_batch(ids[:N_BATCH], [e for e in emb[:N_BATCH]], meta[:N_BATCH])
    for i in range(N_BATCH, N_BATCH + N_SINGLE):
        db.store_embedding(ids[i], emb[i], meta[i])
    rec["raw_vector_before_search"] = np.asarray(db.get_vector(ids[5]), dtype=np.float64).tolist()
    specs = query_specs()
    for i, kw in enumerate(specs):
        out = db.find_most_similar(queries[i], **kw)
        rec["queries_before"].append({"ids": [repr(x) for x in out[0]], "dist": [float(x) for x in out[1]],
                                      "n_meta": len(out[2])})
    for uid in deletes:
        if sharded:
            db.delete_embeddings_batch(uid)
        else:
            db.delete_embedding(uid)
    for i, kw in enumerate(specs):
        out = db.find_most_similar(queries[i], **kw)
        rec["queries_after"].append({"ids": [repr(x) for x in out[0]], "dist": [float(x) for x in out[1]],
                                     "n_meta": len(out[2])})
    rec["inverse_id_map"] = sorted((repr(k), int(v)) for k, v in db.inverse_id_map.items())
    rec["n_metadata"] = len(db.metadata)
    rec["n_embeddings"] = len(db.embeddings)
    if not sharded:
        rec["id_map"] = sorted((int(k), repr(v)) for k, v in db.id_map.items())
    return rec
