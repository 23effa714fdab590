# One-GPU validation of the tree (run under gpurun): the whole -m gpu suite, memcheck + synccheck + initcheck on the two
# small end-to-end scripts, and the k = 10 / 100 fp32 vs int8-shadow timing at the C2 shape.
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for tool in memcheck synccheck initcheck; do
  for script in small tails; do
    timeout 420 compute-sanitizer --tool $tool --error-exitcode 7 python tools/sanitize_$script.py > gpurun_out/sanitize_${script}_$tool.log 2>&1
    echo "$tool $script rc=$? $(grep 'ERROR SUMMARY' gpurun_out/sanitize_${script}_$tool.log | head -1) $(grep -c 'sanitize_.* ok' gpurun_out/sanitize_${script}_$tool.log)"
  done
done
python - <<PY
import sys, json, numpy as np, torch
sys.path.insert(0, ".")
import minivectordb_b200 as mv
n, d = 1_000_000, 384
eng = mv.FlatIPEngine(d, capacity_hint=n); eng.add_synthetic(1234, 0, n, 0, True); ws = eng.workspace()
q = torch.randn(64, d, device="cuda"); q = q / q.norm(dim=1, keepdim=True)
st = torch.cuda.current_stream().cuda_stream
for k in (10, 100):
    D = torch.empty(64, k, device="cuda"); I = torch.empty(64, k, dtype=torch.int64, device="cuda")
    for sh in (0, 1):
        eng.set_option("scan_shadow", sh)
        def go(i): eng.search_device(ws, q[i:i+1].data_ptr(), 1, k, D[i:i+1].data_ptr(), I[i:i+1].data_ptr(), stream=st)
        for i in range(8): go(i)
        torch.cuda.synchronize(); ts = []
        for rep in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(64): go(i)
            e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) / 64 * 1e3)
        print("C2 shape k", k, "shadow", sh, round(float(np.median(ts)), 1), "us")
PY
