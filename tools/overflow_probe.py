"""How often does a small batch overflow its candidate lists (and fall back to the scan)?  Launch counts
and time per batch for several random query sets."""
import os, sys, json
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import minivectordb_b200 as mv
from minivectordb_b200 import _native
n, d, k = 1_000_000, 384, 10
eng = mv.FlatIPEngine(d); eng.add_synthetic(1234, 0, n, 0, True); ws = eng.workspace()
eng.set_option("batch_cost_model", 0)
st = torch.cuda.current_stream().cuda_stream
for nq in (5, 8, 12, 16, 32):
    for trial in range(4):
        torch.manual_seed(100 * nq + trial)
        q = torch.randn(nq, d, device="cuda"); q = q / q.norm(dim=1, keepdim=True)
        D = torch.empty(nq, k, device="cuda"); I = torch.empty(nq, k, dtype=torch.int64, device="cuda")
        eng.search_device(ws, q.data_ptr(), nq, k, D.data_ptr(), I.data_ptr(), stream=st); torch.cuda.synchronize()
        l0 = _native.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); eng.search_device(ws, q.data_ptr(), nq, k, D.data_ptr(), I.data_ptr(), stream=st); e1.record(); torch.cuda.synchronize()
        print(json.dumps(dict(nq=nq, trial=trial, ms=round(e0.elapsed_time(e1), 3), launches=_native.launch_count() - l0)), flush=True)
