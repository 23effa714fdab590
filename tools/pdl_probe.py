"""Programmatic dependent launch (option pdl): device time per search for back-to-back launches on one
stream with and without it, and bit-equality of every result (each query writes its own output slot
and, separately, all queries reuse ONE output slot to exercise the write-after-write ordering)."""
import os, sys, json
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import minivectordb_b200 as mv
out = []
for n, d, k, masked in ((100_000, 512, 10, False), (1_000_000, 384, 10, True), (1_000_000, 384, 100, False), (4736, 512, 10, False)):
    eng = mv.FlatIPEngine(d); eng.add_synthetic(1234, 0, n, 0, True); ws = eng.workspace()
    nbytes = n * eng.device_view()[1] * 4
    NQ = 128
    qs = torch.randn(NQ, d, device="cuda")
    D = torch.empty(NQ, k, device="cuda"); I = torch.empty(NQ, k, dtype=torch.int64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    mptr, mrows = 0, 0
    if masked:
        m = torch.randint(-2**31, 2**31 - 1, ((n + 31) // 32,), dtype=torch.int32, device="cuda")
        if n % 32: m[-1] &= (1 << (n % 32)) - 1
        mptr, mrows = m.data_ptr(), n
    def go(i, slot=None):
        s = i if slot is None else slot
        eng.search_device(ws, qs[i:i+1].data_ptr(), 1, k, D[s:s+1].data_ptr(), I[s:s+1].data_ptr(), mptr, mrows, stream=st)
    ref = None
    for pdl in (0, 1, 0, 1):
        eng.set_option("pdl", pdl)
        for i in range(10): go(i)
        torch.cuda.synchronize()
        ts = []
        for rep in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(NQ): go(i)
            e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) / NQ * 1e-3)
        t = float(np.median(ts))
        cur = (D.clone(), I.clone())
        if ref is None: ref = cur
        same = bool(torch.equal(ref[0], cur[0]) and torch.equal(ref[1], cur[1]))
        # write-after-write: every query into slot 0, the last one must win
        for i in range(NQ): go(i, 0)
        torch.cuda.synchronize()
        waw = bool(torch.equal(D[0], ref[0][NQ - 1]) and torch.equal(I[0], ref[1][NQ - 1]))
        go(0)
        torch.cuda.synchronize()
        rec = dict(n=n, d=d, k=k, masked=masked, pdl=pdl, us=round(t * 1e6, 1), eff_GBs=round(nbytes / t / 1e9), identical=same, waw_ok=waw)
        out.append(rec); print(json.dumps(rec), flush=True)
    del ws; eng.close()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/pdl_probe.json", "w"), indent=1)
