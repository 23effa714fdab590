"""A dozen int8 shadow searches on the C4 shard (12.5M x 512) -- the command `ncu -k regex:scan_i8` profiles."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import minivectordb_b200 as mv
from minivectordb_b200 import synth
n, d = int(os.environ.get("ROWS", 12_500_000)), 512
eng = mv.FlatIPEngine(d, capacity_hint=n)
eng.add_synthetic(1234, 0, n, 0, True)
eng.set_option("scan_shadow", 1)
ws = eng.workspace()
q = synth.synth_rows(4321, 0, 12, d); q /= np.linalg.norm(q, axis=1, keepdims=True)
qd = torch.from_numpy(np.ascontiguousarray(q, dtype=np.float32)).cuda()
D = torch.empty(1, 10, device="cuda"); I = torch.empty(1, 10, dtype=torch.int64, device="cuda")
for i in range(12):
    eng.search_device(ws, qd[i:i + 1].data_ptr(), 1, 10, D.data_ptr(), I.data_ptr(), stream=torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
print(ws.shadow_counters())
