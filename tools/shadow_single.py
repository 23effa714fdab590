"""Opt-in experiment: a SINGLE query through the bf16 shadow + exact re-scoring (half the bytes of the fp32 scan)."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import minivectordb_b200 as mv
n, d, k = int(sys.argv[1]), int(sys.argv[2]), 10
eng = mv.FlatIPEngine(d); eng.add_synthetic(1234, 0, n, 0, True); ws = eng.workspace()
q = torch.randn(1, d, device="cuda"); q = q / q.norm(dim=1, keepdim=True)
D = torch.empty(1, k, device="cuda"); I = torch.empty(1, k, dtype=torch.int64, device="cuda")
st = torch.cuda.current_stream().cuda_stream
res = {}
for name, opts in (("fp32_scan", dict(batch_mode=0)), ("bf16_shadow_exact", dict(batch_mode=1, batch_min_nq=1, batch_cost_model=0))):
    for o, v in opts.items(): eng.set_option(o, v)
    for _ in range(3): eng.search_device(ws, q.data_ptr(), 1, k, D.data_ptr(), I.data_ptr(), stream=st)
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); eng.search_device(ws, q.data_ptr(), 1, k, D.data_ptr(), I.data_ptr(), stream=st); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    res[name] = dict(ms=sorted(ts)[5], ids=I.cpu().tolist(), dist=D.cpu().tolist())
print(json.dumps(dict(n=n, d=d, fp32_scan_ms=res["fp32_scan"]["ms"], shadow_ms=res["bf16_shadow_exact"]["ms"],
                      identical=res["fp32_scan"]["ids"] == res["bf16_shadow_exact"]["ids"] and res["fp32_scan"]["dist"] == res["bf16_shadow_exact"]["dist"])))
