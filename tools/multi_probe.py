import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import minivectordb_b200 as mv
for n, d in ((1_000_000, 384), (1_000_000, 768), (500_000, 1024)):
    eng = mv.FlatIPEngine(d); eng.add_synthetic(1234, 0, n, 0, True); ws = eng.workspace()
    eng.set_option("batch_mode", 0)
    nbytes = n * eng.device_view()[1] * 4
    st = torch.cuda.current_stream().cuda_stream
    for nq in (2, 4, 8):
        q = torch.randn(nq, d, device="cuda"); D = torch.empty(nq, 10, device="cuda"); I = torch.empty(nq, 10, dtype=torch.int64, device="cuda")
        for variant, cws in ((1, (0, 2, 3, 4, 5, 6, 8)), (2, (0,))):
            eng.set_option("scan_variant", variant)
            for cw in cws:
                eng.set_option("consumer_warps", cw)
                for _ in range(3): eng.search_device(ws, q.data_ptr(), nq, 10, D.data_ptr(), I.data_ptr(), stream=st)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(10): eng.search_device(ws, q.data_ptr(), nq, 10, D.data_ptr(), I.data_ptr(), stream=st)
                e1.record(); torch.cuda.synchronize()
                t = e0.elapsed_time(e1) / 10 * 1e-3
                print(json.dumps(dict(n=n, d=d, nq=nq, variant=variant, cw=cw, us=t * 1e6, frac=nbytes / t / 1e9 / 6452.8)), flush=True)
    del ws; eng.close()
