"""Host-to-host latency of mvdb_index_search for ONE query (the e2e leg of bench.py): no filter,
a packed filter uploaded with the call, and a device-resident mask handle."""
import os, sys, json, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import minivectordb_b200 as mv
from minivectordb_b200 import synth
out = []
for n, d, k in ((100_000, 512, 10), (1_000_000, 384, 10), (1_000_000, 384, 100)):
    eng = mv.FlatIPEngine(d); eng.add_synthetic(1234, 0, n, 0, True)
    q = synth.synth_rows(7, 0, 256, d).astype(np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    adm = synth.synth_mask(5, n, 0.5); packed = mv.pack_mask(adm)
    h = eng.mask_handle(adm)
    ref = {}
    for mode in ("none", "bytes", "handle"):
        for zc in (0,):
            def go(i):
                if mode == "none": return eng.search(q[i:i + 1], k)
                if mode == "bytes": return eng.search(q[i:i + 1], k, mask=packed, mask_rows=n)
                return eng.search(q[i:i + 1], k, mask=h)
            for i in range(20): go(i)
            lat = []; res = []
            for i in range(200):
                a = time.perf_counter(); r = go(i); lat.append(time.perf_counter() - a); res.append(r)
            key = mode
            if key not in ref: ref[key] = res
            same = all(np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) for a, b in zip(ref[key], res))
            rec = dict(n=n, d=d, k=k, mask=mode, p50_us=round(float(np.median(lat)) * 1e6, 1),
                       mean_us=round(float(np.mean(lat)) * 1e6, 1), identical=same)
            out.append(rec); print(json.dumps(rec), flush=True)
    eng.close()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/e2e_probe.json", "w"), indent=1)
