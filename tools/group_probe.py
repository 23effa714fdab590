"""Single-process multi-GPU probe (run with `gpurun --gpus N`): the shard group behind
ShardedVectorDatabase(devices=[0..N-1]) on BASELINE config 4 -- 12.5M x 512 fp32 rows per GPU (N = 8:
100M x 512), one query at a time through mvdb_group_search (host buffers in, host results out).
Reports p50 / mean latency of the C-ABI call, parity of 4 queries against the streamed oracle, clocks, and the
Python-level find_most_similar overhead measured on a small database through the drop-in class."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import minivectordb_b200 as mv  # noqa: E402
from minivectordb_b200 import _native, synth  # noqa: E402
from oracle import oracle as O  # noqa: E402

ndev = _native.device_count()
rows = int(os.environ.get("ROWS", 12_500_000))
d, k, steps = 512, 10, int(os.environ.get("STEPS", 200))
engines = []
t0 = time.time()
for s in range(ndev):
    e = mv.FlatIPEngine(d, device=s, capacity_hint=rows)
    e.add_synthetic(bench.SEED_DB, s * rows, rows, 0, True)
    engines.append(e)
grp = mv.ShardGroup(engines)
fill_s = time.time() - t0
q = synth.synth_rows(bench.SEED_Q, 0, steps + 10, d)
q /= np.linalg.norm(q, axis=1, keepdims=True)
q = np.ascontiguousarray(q, dtype=np.float32)
for i in range(10):
    grp.search(q[i:i + 1], k)
clocks = bench.ClockSampler(0)
clocks.start()
lat = []
t0 = time.perf_counter()
for i in range(steps):
    a = time.perf_counter()
    D, S, R = grp.search(q[10 + i:11 + i], k)
    lat.append(time.perf_counter() - a)
total = time.perf_counter() - t0
clk = clocks.stop()
# parity: 4 queries against the oracle streamed over every shard
nq = 4
res = [grp.search(q[10 + i:11 + i], k) for i in range(nq)]
Dg = np.concatenate([r[0] for r in res])
Ig = np.concatenate([np.where(r[1] >= 0, r[1] * rows + r[2], -1) for r in res])
buf = np.empty((1 << 20, d), dtype=np.float32)


def chunk_of(shard):
    def chunk(r0, m):
        x = O.synth_rows(bench.SEED_DB, shard * rows + r0, m, d, out=buf)
        O.normalize_L2(x)
        return x
    return chunk


t1 = time.time()
parts = [O.search_streamed(chunk_of(s), rows, q[10:10 + nq], k, row_offset=s * rows, nthreads=nq) for s in range(ndev)]
Dr, Ir = O.merge_topk_lists(parts, k)


def fetch(label):
    x = O.synth_rows(bench.SEED_DB, int(label), 1, d)
    O.normalize_L2(x)
    return x[0]


rep = O.classify_parity_lazy(fetch, d, q[10:10 + nq], Ig, Dg, Ir, Dr)
oracle_s = time.time() - t1
peak = bench.measured_peaks()[0]
p50 = float(np.median(lat))
out = {"what": "mvdb_group_search (one process, one host thread, fused NVLink exchange), host buffers in/out",
       "n_gpus": ndev, "rows_per_gpu": rows, "rows_total": rows * ndev, "dim": d, "k": k, "steps": steps,
       "p50_latency_ms": p50 * 1e3, "mean_latency_ms": total / steps * 1e3, "qps": steps / total,
       "aggregate_gbs_at_p50": rows * ndev * d * 4 / p50 / 1e9, "aggregate_peak_gbs": peak * ndev,
       "frac_of_aggregate_measured_peak_at_p50": rows * d * 4 / p50 / 1e9 / peak,
       "target_ms_80pct": rows * d * 4 / (0.8 * peak * 1e9) * 1e3,
       "clocks": clk, "fill_s": fill_s,
       "parity": {"ok": bool(rep["ok"]), "queries": nq, "rows_checked": rows * ndev, "ids_equal": rep["id_equal"],
                  "positions": rep["positions"], "exact_ties": rep["exact_tie"], "near_ties": rep["near_tie"],
                  "real_errors": rep["real_error"], "max_rel_err": rep["max_rel_err"], "oracle_seconds": round(oracle_s, 1)}}
grp.close()
for e in engines:
    e.close()

# Python-level overhead of the drop-in class on top of the group call (small database, same code path)
import tempfile  # noqa: E402
from minivectordb_b200 import ShardedVectorDatabase  # noqa: E402
with tempfile.TemporaryDirectory() as tmp:
    db = ShardedVectorDatabase(storage_dir=tmp, devices=list(range(ndev)), persist=False)
    m = 20000
    emb = synth.synth_rows(5, 0, m, d)
    db.store_embeddings_batch(list(range(m)), emb, [{"v": i % 10} for i in range(m)])
    db.find_most_similar(emb[0], k=k)
    lat2, lat3 = [], []
    for i in range(200):
        a = time.perf_counter()
        ids, dist, meta = db.find_most_similar(emb[i], k=k)
        lat2.append(time.perf_counter() - a)
        assert ids[0] == i
        a = time.perf_counter()
        db._group.search(emb[i:i + 1], k, normalize=True)
        lat3.append(time.perf_counter() - a)
    out["dropin_class_small_db"] = {"rows": m, "find_most_similar_p50_us": float(np.median(lat2)) * 1e6,
                                    "group_search_p50_us": float(np.median(lat3)) * 1e6,
                                    "python_overhead_us": float(np.median(lat2) - np.median(lat3)) * 1e6}
    db.close()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open(f"gpurun_out/group_probe_n{ndev}.json", "w"), indent=1)
print(json.dumps(out))
