"""CPU-baseline leg (3) of SURVEY.md section 8(d): the UNMODIFIED reference classes (imported from
/root/reference; `faiss` resolved to the oracle's faiss-shaped layer, `thefuzz` stubbed) on BASELINE config 2
-- what a user of the reference actually waits for per query: Python filter evaluation over sets of row
numbers (ref vector_database.py:354-386), gather of the admissible rows into a temporary index (:508-514),
index rebuild after a mutation (:42-47).  Runs in the build container (the reference checkout does not travel);
the result is committed as profiles/r02_reference_python_path.json and quoted by bench.py's cpu_baseline."""
import json
import os
import sys
import time
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
import bench  # noqa: E402

O.install_as_faiss()
fuzz = types.ModuleType("thefuzz.fuzz"); fuzz.partial_ratio = lambda a, b: 0
tf = types.ModuleType("thefuzz"); tf.fuzz = fuzz
sys.modules["thefuzz"], sys.modules["thefuzz.fuzz"] = tf, fuzz
sys.path.insert(0, "/root/reference")
from minivectordb.vector_database import VectorDatabase  # noqa: E402

n = int(os.environ.get("ROWS", 1_000_000))
d, k = 384, 10
x = O.synth_rows(bench.SEED_DB, 0, n, d)
vals = (O.synth_rows(bench.SEED_META, 0, n, 1, O.DIST_UNIFORM)[:, 0] * 100).astype(int)
db = VectorDatabase(storage_file="/tmp/_ref_probe.pkl")
t0 = time.time()
db.store_embeddings_batch(list(range(n)), list(x), [{"value": int(v)} for v in vals])
t_store = time.time() - t0
q = O.synth_rows(bench.SEED_Q, 0, 8, d)
t0 = time.time()
db.find_most_similar(q[0], k=k)          # first search: builds the index (normalises + copies every row)
t_first = time.time() - t0
unf, fil = [], []
for i in range(1, 4):
    t0 = time.time(); db.find_most_similar(q[i], k=k); unf.append(time.time() - t0)
for i in range(4, 7):
    t0 = time.time(); ids, dist, meta = db.find_most_similar(q[i], metadata_filter={"value": {"$gt": 49}}, k=k); fil.append(time.time() - t0)
    assert all(m["value"] > 49 for m in meta)
t0 = time.time()
db.store_embedding("extra", x[0], {"value": 1})
db.find_most_similar(q[7], k=k)          # a search after ONE insert pays np.vstack + the full index rebuild
t_after_insert = time.time() - t0
out = {"what": "unmodified reference VectorDatabase (Python) over the oracle's faiss-shaped layer, single thread, build container",
       "config": f"{n} x {d} fp32, k={k}, metadata {{'value': int U[0,100)}}, filter {{'value': {{'$gt': 49}}}}",
       "cores_available": os.cpu_count(), "store_embeddings_batch_s": round(t_store, 2), "first_search_with_index_build_s": round(t_first, 3),
       "unfiltered_query_s": round(float(np.median(unf)), 4), "filtered_query_s": round(float(np.median(fil)), 3),
       "search_after_one_insert_s": round(t_after_insert, 3),
       "unfiltered_qps": round(1.0 / float(np.median(unf)), 2), "filtered_qps": round(1.0 / float(np.median(fil)), 3)}
json.dump(out, open(os.path.join(ROOT, "profiles", "r02_reference_python_path.json"), "w"), indent=1)
print(json.dumps(out))
