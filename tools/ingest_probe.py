"""Roofline of the ingest kernel (mvdb_index_add_device): n*d*4 bytes read + n*ld*4 bytes written."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import minivectordb_b200 as mv
import bench
peak = bench.measured_peaks()[0]
out = []
for n, d in ((1_000_000, 384), (2_000_000, 512), (1_000_000, 1024), (500_000, 100)):
    x = torch.randn(n, d, device="cuda")
    ts = []
    eng = mv.FlatIPEngine(d, capacity_hint=n)
    eng.add_device(x.data_ptr(), n, True)          # maps (and zero-fills) the physical memory once
    for rep in range(5):
        eng.reset()                                 # keeps the mapping: the timed add only runs the kernel
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        eng.add_device(x.data_ptr(), n, True)      # synchronous: returns when the rows are resident
        ts.append(time.perf_counter() - t0)
    eng.close()
    ld = (d + 3) // 4 * 4
    t = float(np.median(ts))
    rec = dict(n=n, d=d, ms=t * 1e3, gbs=(n * d * 4 + n * ld * 4) / t / 1e9, frac_of_copy_peak=(n * d * 4 + n * ld * 4) / t / 1e9 / peak,
               note="wall time of mvdb_index_add_device into already mapped memory (kernel + bookkeeping + one sync)")
    out.append(rec); print(json.dumps(rec), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/ingest_probe.json", "w"), indent=1)
