"""Fixed cost of one scan launch: device time per launch (back-to-back, CUDA events) for tiny to
mid-size matrices, so that the epilogue/launch share of the small-N configs is visible."""
import os, sys, json
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import minivectordb_b200 as mv
out = []
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for d in (512,):
    for n in (8, 1184, 4736, 18944, 50_000, 100_000, 200_000):
        for k in (1, 10, 100):
            eng = mv.FlatIPEngine(d); eng.add_synthetic(1234, 0, n, 0, True); ws = eng.workspace()
            qs = torch.randn(64, d, device="cuda")
            D = torch.empty(64, k, device="cuda"); I = torch.empty(64, k, dtype=torch.int64, device="cuda")
            st = torch.cuda.current_stream().cuda_stream
            def go(i): eng.search_device(ws, qs[i:i+1].data_ptr(), 1, k, D[i:i+1].data_ptr(), I[i:i+1].data_ptr(), stream=st)
            for i in range(10): go(i)
            torch.cuda.synchronize()
            ts = []
            for rep in range(5):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for i in range(64): go(i)
                e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1) / 64 * 1e3)
            rec = dict(n=n, d=d, k=k, us=round(float(np.median(ts)), 2), stream_us=round(n * d * 4 / 6.4528e6, 2))
            out.append(rec); print(json.dumps(rec), flush=True)
            del ws; eng.close()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/fixed_cost_probe.json", "w"), indent=1)
