"""Launch one scan configuration a few times (for ncu): python tools/one_config.py n d nq k [variant] [iters]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import minivectordb_b200 as mv  # noqa: E402

n, d, nq, k = (int(a) for a in sys.argv[1:5])
variant = int(sys.argv[5]) if len(sys.argv) > 5 else 0
iters = int(sys.argv[6]) if len(sys.argv) > 6 else 6
eng = mv.FlatIPEngine(d)
eng.add_synthetic(1234, 0, n, 0, True)
eng.set_option("scan_variant", variant)
eng.set_option("batch_mode", 0)
ws = eng.workspace()
q = torch.randn(nq, d, device="cuda")
D = torch.empty(nq, k, device="cuda")
I = torch.empty(nq, k, dtype=torch.int64, device="cuda")
for _ in range(iters):
    eng.search_device(ws, q.data_ptr(), nq, k, D.data_ptr(), I.data_ptr(), normalize=True,
                      stream=torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
print("ok", I[0, :3].tolist())
