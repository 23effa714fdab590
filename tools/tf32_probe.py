"""batch_mode 1 (exact) / 2 (bf16) / 3 (tf32): time and recall@k against the exact mode."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import minivectordb_b200 as mv
out = []
for n, d, nq, k in [((1_000_000, 384, 4096, 10), (2_000_000, 1024, 4096, 100), (10_000_000, 1024, 4096, 100))[int(c)] for c in os.environ.get("CASES", "0,1").split(",")]:
    eng = mv.FlatIPEngine(d); eng.add_synthetic(1234, 0, n, 0, True); ws = eng.workspace()
    q = torch.randn(nq, d, device="cuda"); q = q / q.norm(dim=1, keepdim=True)
    D = torch.empty(nq, k, device="cuda"); I = torch.empty(nq, k, dtype=torch.int64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    ref = None
    for mode in (1, 2, 3):
        eng.set_option("batch_mode", mode)
        eng.search_device(ws, q.data_ptr(), nq, k, D.data_ptr(), I.data_ptr(), stream=st); torch.cuda.synchronize()
        ts = []
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); eng.search_device(ws, q.data_ptr(), nq, k, D.data_ptr(), I.data_ptr(), stream=st); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        t = sorted(ts)[1]
        if ref is None: ref = (D.clone(), I.clone())
        Ic, Ir = I.cpu().numpy(), ref[1].cpu().numpy()
        recall = float(sum(len(set(a) & set(b)) for a, b in zip(Ic, Ir)) / Ir.size)
        rec = dict(n=n, d=d, nq=nq, k=k, mode={1: "exact", 2: "bf16", 3: "tf32"}[mode], ms=round(t, 3),
                   tflops=round(2.0 * nq * n * d / t / 1e9, 1), recall_vs_exact=round(recall, 5),
                   max_abs_score_diff=float((D - ref[0]).abs().max()))
        out.append(rec); print(json.dumps(rec), flush=True)
    del ws; eng.close()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/tf32_probe.json", "w"), indent=1)
