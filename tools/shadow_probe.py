"""int8 shadow mode vs fp32 scan: device time per search (back-to-back launches on one stream, CUDA events)
and host-to-host latency through mvdb_index_search, on C1 / C2 / the C4 shard; bytes moved and roofline."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import minivectordb_b200 as mv
import bench
from minivectordb_b200 import synth
peak = bench.measured_peaks()[0]
out = []
for name, n, d in (("C1", 100_000, 512), ("C2", 1_000_000, 384), ("C4 shard", 12_500_000, 512), ("C5 shape", 10_000_000, 768)):
    eng = mv.FlatIPEngine(d, capacity_hint=n)
    eng.add_synthetic(1234, 0, n, 0, True)
    eng.set_option("coalesce", 0)
    ws = eng.workspace()
    q = synth.synth_rows(4321, 0, 72, d); q /= np.linalg.norm(q, axis=1, keepdims=True); q = np.ascontiguousarray(q, dtype=np.float32)
    qd = torch.from_numpy(q).cuda()
    D = torch.empty(64, 10, device="cuda"); I = torch.empty(64, 10, dtype=torch.int64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    rec = dict(config=name, n=n, d=d)
    res = {}
    for mode in (0, 1):
        eng.set_option("scan_shadow", mode)
        def go(i): eng.search_device(ws, qd[i:i+1].data_ptr(), 1, 10, D[i:i+1].data_ptr(), I[i:i+1].data_ptr(), stream=st)
        for i in range(8): go(i)
        torch.cuda.synchronize()
        ts = []
        for rep in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(64): go(i)
            e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) / 64 * 1e3)
        res[mode] = (D.cpu().numpy().copy(), I.cpu().numpy().copy())
        if mode == 1:
            cs = []
            for i in range(16):
                go(i); cs.append(ws.shadow_counters())
            rec["candidates_median"] = int(np.median([c[0] for c in cs])); rec["survivors_median"] = int(np.median([c[1] for c in cs]))
            rec["overflows_of_16"] = int(sum(c[2] for c in cs))
        lat = []
        for i in range(64):
            a = time.perf_counter(); eng.search(q[i:i+1], 10); lat.append(time.perf_counter() - a)
        byts = n * (d * 4) if mode == 0 else n * (((d + 15) // 16) * 16 + 16)
        key = "fp32" if mode == 0 else "int8_shadow"
        rec[key] = dict(device_us=round(float(np.median(ts)), 2), host_p50_us=round(float(np.median(lat)) * 1e6, 2), stream_bytes=byts,
                        gbs=round(byts / (np.median(ts) * 1e-6) / 1e9, 1), frac_of_copy_peak=round(byts / (np.median(ts) * 1e-6) / 1e9 / peak, 3))
    rec["identical"] = bool(np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1]))
    rec["speedup_device"] = round(rec["fp32"]["device_us"] / rec["int8_shadow"]["device_us"], 2)
    out.append(rec); print(json.dumps(rec), flush=True)
    ws.close(); eng.close()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/shadow_probe.json", "w"), indent=1)
