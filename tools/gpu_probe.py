"""Device-side timing sweep of the scan variants (run under gpurun).
Writes gpurun_out/probe.json.  Timing: CUDA events on the launch stream,
warm-up 5, matrix >> L2 (or L2 flushed when it is not)."""
import ctypes
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import minivectordb_b200 as mv  # noqa: E402
from minivectordb_b200 import _native as N  # noqa: E402

PEAK = 6452.8
try:
    PEAK = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]
except Exception:
    pass

out = []
torch.cuda.init()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def time_search(eng, ws, nq, k, mask_ptr, mask_rows, iters=20, flush_l2=False):
    """Per-launch device time with the launches enqueued BACK TO BACK (the host
    enqueue cost of python+ctypes, ~15-25 us, would otherwise sit between the two
    events of a single launch).  Cold numbers: (flush + search) pairs minus the
    flush alone, both back to back."""
    d = eng.d
    q = torch.randn(nq, d, device="cuda")
    q = q / q.norm(dim=1, keepdim=True)
    D = torch.empty(nq, k, device="cuda")
    I = torch.empty(nq, k, dtype=torch.int64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream

    def run(n_it, with_search, with_flush):
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(n_it):
            if with_flush:
                flush.fill_(1)
            if with_search:
                eng.search_device(ws, q.data_ptr(), nq, k, D.data_ptr(), I.data_ptr(), mask_ptr, mask_rows, stream=st)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e-3 / n_it

    run(5, True, False)
    reps = []
    for _ in range(5):
        if flush_l2:
            reps.append(run(iters, True, True) - run(iters, False, True))
        else:
            reps.append(run(iters, True, False))
    reps.sort()
    return reps[len(reps) // 2], reps[0]


configs = [(1_000_000, 384), (100_000, 512), (2_000_000, 512), (1_000_000, 768), (1_000_000, 1024)]
mode = sys.argv[1] if len(sys.argv) > 1 else "full"
if mode == "quick":
    configs = configs[:2]
if mode == "latency":
    # fixed-cost anatomy: 1 tile per CTA, 10 tiles per CTA, config 1 (100k x 512), cold (L2 flushed) and warm
    for n, d in [(1184, 512), (11840, 512), (100_000, 512), (100_000, 384)]:
        eng = mv.FlatIPEngine(d)
        eng.add_synthetic(1234, 0, n, 0, True)
        ws = eng.workspace()
        nbytes = n * eng.device_view()[1] * 4
        for variant in (1, 2):
            eng.set_option("scan_variant", variant)
            for k in (10, 100):
                for cold in (True, False):
                    med, best = time_search(eng, ws, 1, k, 0, n, iters=30, flush_l2=cold)
                    rec = dict(mode="latency", n=n, d=d, variant=variant, k=k, cold=cold, med_us=med * 1e6,
                               best_us=best * 1e6, gbs=nbytes / med / 1e9)
                    out.append(rec)
                    print(json.dumps(rec), flush=True)
        del ws
        eng.close()
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/probe_latency.json", "w"), indent=1)
    sys.exit(0)
for n, d in configs:
    eng = mv.FlatIPEngine(d)
    t0 = time.time()
    eng.add_synthetic(1234, 0, n, 0, True)
    t_add = time.time() - t0
    ws = eng.workspace()
    mask = torch.randint(0, 2**31 - 1, ((n + 31) // 32,), dtype=torch.int32, device="cuda")
    nbytes = n * eng.device_view()[1] * 4
    small = nbytes < (300 << 20)
    for variant, cw_list in ((1, (2, 4, 8)), (2, (0,))):
        for cw in cw_list:
            for grid_mult in ((1,) if variant == 1 else (0, 2, 4)):
                eng.set_option("scan_variant", variant)
                eng.set_option("consumer_warps", cw)
                eng.set_option("grid_ctas", 0 if variant == 1 else 148 * grid_mult)
                for nq, k, use_mask in ((1, 10, False), (1, 10, True), (1, 100, False), (4, 10, False), (8, 10, False)):
                    try:
                        med, best = time_search(eng, ws, nq, k, mask.data_ptr() if use_mask else 0, n, flush_l2=small)
                    except Exception as e:  # noqa: BLE001
                        print("FAIL", n, d, variant, cw, nq, k, e)
                        continue
                    rec = dict(n=n, d=d, variant=variant, cw=cw, grid_mult=grid_mult, nq=nq, k=k, mask=use_mask,
                               med_us=med * 1e6, best_us=best * 1e6, gbs=nbytes / med / 1e9,
                               frac=nbytes / med / 1e9 / PEAK, t_add=t_add)
                    out.append(rec)
                    print(json.dumps(rec), flush=True)
    del ws
    eng.close()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/probe.json", "w"), indent=1)
