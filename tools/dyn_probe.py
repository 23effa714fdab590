"""Static round-robin vs dynamic (counter-claimed) tile schedule of the TMA scan: device time per
launch (back-to-back launches, CUDA events) for several dynamic percentages, and bit-equality of
the results against the static schedule."""
import os, sys, json
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import minivectordb_b200 as mv
out = []
PCTS = [int(x) for x in os.environ.get("PCTS", "0,15,30,50,100").split(",")]
CASES = ((100_000, 512, 10, False, 1), (1_000_000, 384, 10, True, 1), (1_000_000, 384, 100, False, 1),
         (4_000_000, 768, 10, False, 1), (1_000_003, 100, 10, True, 1), (1_000_000, 384, 10, True, 2),
         (1_000_000, 384, 10, False, 4), (300_000, 2048, 10, True, 1))
for n, d, k, masked, nq in CASES:
    eng = mv.FlatIPEngine(d); eng.add_synthetic(1234, 0, n, 0, True); ws = eng.workspace()
    eng.set_option("batch_mode", 0)
    nbytes = n * eng.device_view()[1] * 4
    qs = torch.randn(64 * nq, d, device="cuda")
    D = torch.empty(64 * nq, k, device="cuda"); I = torch.empty(64 * nq, k, dtype=torch.int64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    mptr, mrows = 0, 0
    if masked:
        m = torch.randint(-2**31, 2**31 - 1, ((n + 31) // 32,), dtype=torch.int32, device="cuda")
        if n % 32: m[-1] &= (1 << (n % 32)) - 1
        mptr, mrows = m.data_ptr(), n
    def go(i):
        a, b = i * nq, (i + 1) * nq
        eng.search_device(ws, qs[a:b].data_ptr(), nq, k, D[a:b].data_ptr(), I[a:b].data_ptr(), mptr, mrows, stream=st)
    ref = None
    for dyn in PCTS:
        eng.set_option("dyn_tiles", dyn)
        for i in range(10): go(i)
        torch.cuda.synchronize()
        ts = []
        for rep in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(64): go(i)
            e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) / 64 * 1e-3)
        t = float(np.median(ts))
        cur = (D.clone(), I.clone())
        if ref is None: ref = cur
        same = bool(torch.equal(ref[0], cur[0]) and torch.equal(ref[1], cur[1]))
        rec = dict(n=n, d=d, k=k, nq=nq, masked=masked, dyn=dyn, us=round(t * 1e6, 1), eff_GBs=round(nbytes / t / 1e9), identical=same)
        out.append(rec); print(json.dumps(rec), flush=True)
    del ws; eng.close()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/dyn_probe.json", "w"), indent=1)
