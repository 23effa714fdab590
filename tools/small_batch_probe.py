"""Small query batches (one query block) on the tensor-core path: short A tile + deeper ring
(option gemm_short_a) vs the fixed 128-row A tile.  Device time per batch, effective bf16 bytes/s."""
import os, sys, json
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import minivectordb_b200 as mv
out = []
for n, d in ((10_000_000, 768), (1_000_000, 384)):
    eng = mv.FlatIPEngine(d); eng.add_synthetic(1234, 0, n, 0, True); ws = eng.workspace()
    eng.set_option("batch_cost_model", 0)
    k = 10
    st = torch.cuda.current_stream().cuda_stream
    for nq in (8, 32, 64, 128):
        q = torch.randn(nq, d, device="cuda"); q = q / q.norm(dim=1, keepdim=True)
        D = torch.empty(nq, k, device="cuda"); I = torch.empty(nq, k, dtype=torch.int64, device="cuda")
        ref = None
        for short in (0, 1, 0, 1):
            eng.set_option("gemm_short_a", short)
            for _ in range(2): eng.search_device(ws, q.data_ptr(), nq, k, D.data_ptr(), I.data_ptr(), stream=st)
            torch.cuda.synchronize()
            ts = []
            for _ in range(5):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); eng.search_device(ws, q.data_ptr(), nq, k, D.data_ptr(), I.data_ptr(), stream=st); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            t = float(np.median(ts))
            cur = (D.clone(), I.clone())
            if ref is None: ref = cur
            same = bool(torch.equal(ref[0], cur[0]) and torch.equal(ref[1], cur[1]))
            rec = dict(n=n, d=d, nq=nq, short_a=short, ms=round(t, 3), shadow_TBs=round(n * d * 2 / t / 1e9, 2), identical=same)
            out.append(rec); print(json.dumps(rec), flush=True)
    del ws; eng.close()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/small_batch_probe.json", "w"), indent=1)
