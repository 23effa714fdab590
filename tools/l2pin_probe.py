import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import minivectordb_b200 as mv
for n, d in ((50_000, 512), (100_000, 512), (200_000, 512), (1_000_000, 384)):
    eng = mv.FlatIPEngine(d); eng.add_synthetic(1234, 0, n, 0, True); ws = eng.workspace()
    nbytes = n * eng.device_view()[1] * 4
    qs = torch.randn(64, d, device="cuda")
    D = torch.empty(1, 10, device="cuda"); I = torch.empty(1, 10, dtype=torch.int64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for pin in (0, 32, 64, 96, 120):
        eng.set_option("l2_pin_mb", pin)
        for i in range(10): eng.search_device(ws, qs[i:i+1].data_ptr(), 1, 10, D.data_ptr(), I.data_ptr(), stream=st)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(50): eng.search_device(ws, qs[i:i+1].data_ptr(), 1, 10, D.data_ptr(), I.data_ptr(), stream=st)
        e1.record(); torch.cuda.synchronize()
        t = e0.elapsed_time(e1) / 50 * 1e-3
        print(json.dumps(dict(n=n, d=d, MB=nbytes >> 20, pin_mb=pin, us=round(t * 1e6, 1), eff_GBs=round(nbytes / t / 1e9))), flush=True)
    del ws; eng.close()
