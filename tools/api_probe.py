"""End-user latency of the drop-in Python API (VectorDatabase.find_most_similar) on the GPU box.
  python tools/api_probe.py [rows] [dim]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from minivectordb_b200 import VectorDatabase, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 384
db = VectorDatabase(storage_file="/tmp/api_probe_none.pkl")
step = 100_000
rng = np.random.default_rng(0)
t_store = 0.0      # time inside store_embeddings_batch only (generating the synthetic rows is not the database's work)
for a in range(0, n, step):
    m = min(step, n - a)
    emb = synth.synth_rows(1234, a, m, d)
    vals = rng.integers(0, 100, m)
    ids, rows, metas = list(range(a, a + m)), list(emb), [{"value": int(v), "tag": f"t{int(v) % 16}"} for v in vals]
    t0 = time.perf_counter()
    db.store_embeddings_batch(ids, rows, metas)
    t_store += time.perf_counter() - t0
q = synth.synth_rows(4321, 0, 300, d)
t0 = time.perf_counter()
db.find_most_similar(q[0], k=10)
t_first = time.perf_counter() - t0


def lat(fn, reps=100):
    ts = []
    for i in range(reps):
        a = time.perf_counter()
        fn(i)
        ts.append(time.perf_counter() - a)
    return float(np.median(ts) * 1e3), float(np.percentile(ts, 95) * 1e3)


out = dict(rows=n, dim=d, store_s=t_store, first_search_flush_s=t_first)
out["unfiltered_ms_p50_p95"] = lat(lambda i: db.find_most_similar(q[i], k=10))
t0 = time.perf_counter()
db.find_most_similar(q[0], k=10, metadata_filter={"value": {"$gt": 49}})
out["filtered_first_evaluation_ms"] = (time.perf_counter() - t0) * 1e3
out["filtered_repeated_ms_p50_p95"] = lat(lambda i: db.find_most_similar(q[i], k=10, metadata_filter={"value": {"$gt": 49}}))
out["filtered_new_filter_each_time_ms_p50_p95"] = lat(
    lambda i: db.find_most_similar(q[i], k=10, metadata_filter={"value": {"$gt": i % 90}}), reps=30)
out["or_exclude_ms_p50_p95"] = lat(lambda i: db.find_most_similar(q[i], k=10, or_filters=[{"tag": "t3"}, {"value": {"$lt": 5}}],
                                                              exclude_filter={"tag": "t7"}), reps=30)
print(json.dumps(out))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/api_probe.json", "w"), indent=1)
