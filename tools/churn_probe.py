"""BASELINE config 5: insert/delete churn with concurrent multithreaded queries and metadata
filters (tombstone + mask path), at the engine (C ABI) level.  Run under gpurun.
  python tools/churn_probe.py [rows] [dim] [seconds]"""
import json
import os
import sys
import threading
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import minivectordb_b200 as mv  # noqa: E402
from minivectordb_b200 import synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 768
secs = float(sys.argv[3]) if len(sys.argv) > 3 else 20.0
k = 10
out = []
# the probe's own threads are Python: with the default 5 ms GIL switch interval a query thread that comes back from
# the C ABI can wait milliseconds for the interpreter while an insert thread builds its next block
sys.setswitchinterval(float(os.environ.get("SWITCH_INTERVAL", "0.0002")))
# lines: (coalesce, query threads, filters as mask handles, coalesce_leaders [0 = auto][, scan_shadow]); override with
# CONFIGS="1:8:1:2,1:8:1:0" ; a fifth field 1 runs the line in the opt-in int8 shadow mode
lines = ((0, 8, 0, 0), (1, 8, 0, 0), (1, 8, 1, 0), (1, 32, 1, 0), (1, 64, 1, 0), (0, 8, 1, 0, 1), (0, 2, 1, 0, 1))
if os.environ.get("CONFIGS"):
    lines = tuple(tuple(int(x) for x in c.split(":")) for c in os.environ["CONFIGS"].split(","))
for line in lines:
    coalesce, nthreads, handles, leaders = line[:4]
    shadow = line[4] if len(line) > 4 else 0
    eng = mv.FlatIPEngine(d, capacity_hint=n + 2_000_000)
    eng.add_synthetic(1234, 0, n, 0, True)
    eng.set_option("coalesce", coalesce)
    eng.set_option("scan_shadow", shadow)
    eng.set_option("coalesce_leaders", leaders)
    masks = [mv.pack_mask(synth.synth_mask(100 + i, n, 0.5)) for i in range(4)]
    if handles:
        masks_h = [eng.mask_handle(synth.synth_mask(100 + i, n, 0.5)) for i in range(4)]
    stop = threading.Event()
    lat = [[] for _ in range(nthreads)]
    counts = dict(ins=0, dele=0)
    errs = []

    def querier(t):
        rng = np.random.default_rng(t)
        try:
            while not stop.is_set():
                q = rng.standard_normal((1, d)).astype(np.float32)
                a = time.perf_counter()
                if rng.random() < 0.5:   # 50 % of the queries carry a metadata filter
                    if handles:
                        eng.search(q, k, mask=masks_h[int(rng.integers(0, 4))], normalize=True)
                    else:
                        eng.search(q, k, mask=masks[int(rng.integers(0, 4))], mask_rows=n, normalize=True)
                else:
                    eng.search(q, k, normalize=True)
                lat[t].append(time.perf_counter() - a)
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    def inserter(t):
        rng = np.random.default_rng(1000 + t)
        try:
            while not stop.is_set():
                eng.add(rng.standard_normal((16, d)).astype(np.float32), normalize=True)
                counts["ins"] += 16
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    def deleter():
        nxt = 0
        try:
            while not stop.is_set() and nxt + 16 < n:
                eng.remove_rows(np.arange(nxt, nxt + 16))
                nxt += 16
                counts["dele"] += 16
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    ts = [threading.Thread(target=querier, args=(t,)) for t in range(nthreads)]
    ts += [threading.Thread(target=inserter, args=(t,)) for t in range(int(os.environ.get("INSERTERS", "2")))]
    ts += [threading.Thread(target=deleter)] * int(os.environ.get("DELETER", "1"))
    t0 = time.perf_counter()
    [t.start() for t in ts]
    time.sleep(secs)
    stop.set()
    [t.join() for t in ts]
    dt = time.perf_counter() - t0
    all_lat = np.concatenate([np.asarray(x) for x in lat if x])
    rec = dict(rows=n, dim=d, k=k, query_threads=nthreads, coalesce=coalesce, mask_handles=handles, leaders=leaders, scan_shadow=shadow, seconds=dt, queries=int(all_lat.size),
               qps=all_lat.size / dt, p50_ms=float(np.median(all_lat) * 1e3), p99_ms=float(np.percentile(all_lat, 99) * 1e3),
               inserted=counts["ins"], deleted=counts["dele"], ntotal=eng.ntotal, nlive=eng.nlive, errors=len(errs),
               scan_bytes=n * eng.device_view()[1] * 4)
    rec["effective_scan_GBs"] = rec["qps"] * rec["scan_bytes"] / 1e9
    out.append(rec)
    print(json.dumps(rec), flush=True)
    assert eng.ntotal == n + counts["ins"] and eng.nlive == n + counts["ins"] - counts["dele"]
    eng.close()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/churn_probe.json", "w"), indent=1)
