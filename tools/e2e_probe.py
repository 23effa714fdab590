"""Host-to-host latency of mvdb_index_search for ONE query (the e2e leg of bench.py) under the pieces of the
"host_path" option: 0 = copy-engine staging both ways (round 1), 1 = results written straight to pinned host memory,
3 = + inputs pulled by a grid the scan depends on programmatically (default).
Cases: no filter, a packed filter uploaded with the call, a device-resident mask handle; fp32 scan and int8 shadow.
Every mode's answers must be identical; the first mode is also checked against the CPU oracle on 8 queries."""
import os, sys, json, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import minivectordb_b200 as mv
from minivectordb_b200 import synth
from oracle import oracle as orc
out = []
CASES = ((100_000, 512, 10, 0), (1_000_000, 384, 10, 0), (1_000_000, 384, 100, 0), (1_000_000, 384, 10, 1))
for n, d, k, shadow in CASES:
    eng = mv.FlatIPEngine(d)
    if shadow: eng.set_option("scan_shadow", 1)
    eng.add_synthetic(1234, 0, n, 0, True)
    q = synth.synth_rows(7, 0, 256, d).astype(np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    adm = synth.synth_mask(5, n, 0.5); packed = mv.pack_mask(adm)
    h = eng.mask_handle(adm)
    x = eng.reconstruct_n(0, n)
    ref = {}
    for mode in ("none", "bytes", "handle"):
        for hp in (0, 1, 3):
            eng.set_option("host_path", hp)
            def go(i):
                if mode == "none": return eng.search(q[i:i + 1], k)
                if mode == "bytes": return eng.search(q[i:i + 1], k, mask=packed, mask_rows=n)
                return eng.search(q[i:i + 1], k, mask=h)
            for i in range(20): go(i)
            lat = []; res = []
            for i in range(200):
                a = time.perf_counter(); r = go(i); lat.append(time.perf_counter() - a); res.append(r)
            if mode not in ref:
                ref[mode] = res
                rows = None if mode == "none" else np.flatnonzero(adm)
                for i in range(8):
                    Do, Io = orc.search_flat_ip(x if rows is None else x[rows], q[i:i + 1], k)
                    if rows is not None: Io = np.where(Io >= 0, rows[np.maximum(Io, 0)], -1)
                    v = orc.classify_parity(x, q[i:i + 1], res[i][1], res[i][0], Io, Do, admissible=None if rows is None else adm)
                    assert v["real_error"] == 0, (n, d, k, mode, v)
            same = all(np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) for a, b in zip(ref[mode], res))
            rec = dict(n=n, d=d, k=k, shadow=shadow, mask=mode, host_path=hp, p50_us=round(float(np.median(lat)) * 1e6, 1),
                       mean_us=round(float(np.mean(lat)) * 1e6, 1), identical=same)
            out.append(rec); print(json.dumps(rec), flush=True)
            assert same, rec
    eng.close()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/e2e_probe.json", "w"), indent=1)
