import os, sys, ctypes
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import minivectordb_b200 as mv
from minivectordb_b200 import _native as N
# k <= 16 takes the extraction tail (no slot 9): 4 = CTA's k best extracted + written, 10 = this warp's slice of the
# G x k partial keys loaded + extracted, 12 = final extraction over the warps' lists
names = {0: "cta0 consumer start", 1: "cta0 query loaded", 2: "cta0 tiles done", 3: "cta0 (compaction+) barrier", 4: "cta0 CTA merge + partials written",
         5: "cta0 fence+barrier", 6: "cta0 ticket taken", 8: "last: start", 9: "last: partials staged", 10: "last: slice merged",
         11: "last: barrier", 12: "last: final merge done", 13: "last: results written"}
K = int(os.environ.get("K", "10"))
for n, d in ((8, 512), (1184, 512), (100_000, 512), (1_000_000, 384)):
    eng = mv.FlatIPEngine(d); eng.add_synthetic(1234, 0, n, 0, True); ws = eng.workspace()
    eng.set_option("trace", 1)
    q = torch.randn(1, d, device="cuda"); D = torch.empty(1, K, device="cuda"); I = torch.empty(1, K, dtype=torch.int64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(5): eng.search_device(ws, q.data_ptr(), 1, K, D.data_ptr(), I.data_ptr(), stream=st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); eng.search_device(ws, q.data_ptr(), 1, K, D.data_ptr(), I.data_ptr(), stream=st); e1.record(); torch.cuda.synchronize()
    out = np.zeros(16, dtype=np.uint64)
    N.check(N.lib().mvdb_debug_read_trace(eng.handle, out.ctypes.data))
    t0 = int(out[0])
    print(f"--- n={n} d={d}: event time {e0.elapsed_time(e1)*1e3:.1f} us")
    for i in sorted(names):
        if out[i]: print(f"  {names[i]:38s} +{(int(out[i]) - t0)/1e3:8.2f} us")
    del ws; eng.close()
