"""Timing of the tcgen05 batched path (run under gpurun).  Writes gpurun_out/batched_probe.json."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import minivectordb_b200 as mv  # noqa: E402

PEAK_TF = 1668.4
try:
    PEAK_TF = json.load(open("MEASURED_PEAKS.json"))["bf16_tflops"]
except Exception:
    pass
out = []
configs = [(1_000_000, 384, [64, 256, 1024, 4096], 10), (1_000_000, 1024, [4096], 100), (10_000_000, 1024, [4096], 100)]
if len(sys.argv) > 1 and sys.argv[1] == "small":
    configs = configs[:2]
if len(sys.argv) > 1 and sys.argv[1] == "crossover":
    configs = [(1_000_000, 384, [4, 8, 12, 16, 24, 32, 48], 10), (1_000_000, 1024, [8, 16, 32], 10)]
if len(sys.argv) > 1 and sys.argv[1] == "ncu":
    configs = [(1_000_000, 1024, [4096], 100)]
for n, d, nqs, k in configs:
    eng = mv.FlatIPEngine(d)
    eng.add_synthetic(1234, 0, n, 0, True)
    ws = eng.workspace()
    st = torch.cuda.current_stream().cuda_stream
    for nq in nqs:
        q = torch.randn(nq, d, device="cuda")
        q = q / q.norm(dim=1, keepdim=True)
        D = torch.empty(nq, k, device="cuda")
        I = torch.empty(nq, k, dtype=torch.int64, device="cuda")
        ref = None
        for mode in (1, 2, 0):
            eng.set_option("batch_min_nq", 1)
            eng.set_option("batch_cost_model", 0)
            if mode == 0 and nq > 256:
                continue  # fp32 scan of thousands of queries is only a sanity baseline
            eng.set_option("batch_mode", mode)
            eng.search_device(ws, q.data_ptr(), nq, k, D.data_ptr(), I.data_ptr(), stream=st)  # warm-up (+ shadow build)
            torch.cuda.synchronize()
            ts = []
            for _ in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                eng.search_device(ws, q.data_ptr(), nq, k, D.data_ptr(), I.data_ptr(), stream=st)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1) * 1e-3)
            t = sorted(ts)[1]
            rec = dict(n=n, d=d, nq=nq, k=k, mode={0: "scan_fp32", 1: "tc_exact", 2: "tc_bf16"}[mode], s=t,
                       qps=nq / t, tflops=2.0 * nq * n * d / t / 1e12, frac_bf16_peak=2.0 * nq * n * d / t / 1e12 / PEAK_TF)
            if mode == 1:
                ref = (D.clone(), I.clone())
            elif mode == 2 and ref is not None:
                Ia, Ib = ref[1].cpu().numpy(), I.cpu().numpy()
                rec["recall_vs_exact"] = float(np.mean([len(set(Ia[i]) & set(Ib[i])) / k for i in range(min(nq, 512))]))
            elif mode == 0 and ref is not None:
                rec["ids_equal_exact"] = bool(torch.equal(ref[1], I))
                rec["dist_equal_exact"] = bool(torch.equal(ref[0], D))
            out.append(rec)
            print(json.dumps(rec), flush=True)
        if nq >= 4096 and ref is not None:
            # parity spot check at scale: 8 of the queries through the fp32 scan
            eng.set_option("batch_mode", 0)
            D8 = torch.empty(8, k, device="cuda")
            I8 = torch.empty(8, k, dtype=torch.int64, device="cuda")
            eng.search_device(ws, q[:8].contiguous().data_ptr(), 8, k, D8.data_ptr(), I8.data_ptr(), stream=st)
            torch.cuda.synchronize()
            rec = dict(n=n, d=d, nq=nq, k=k, mode="parity_spot_check_8q", ids_equal=bool(torch.equal(I8, ref[1][:8])),
                       dist_equal=bool(torch.equal(D8, ref[0][:8])))
            out.append(rec)
            print(json.dumps(rec), flush=True)
    del ws
    eng.close()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/batched_probe.json", "w"), indent=1)
