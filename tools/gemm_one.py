import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import minivectordb_b200 as mv
variant = int(sys.argv[1])
n, d, nq, k = 2_000_000, 1024, 4096, 100
eng = mv.FlatIPEngine(d); eng.add_synthetic(1234, 0, n, 0, True); ws = eng.workspace()
eng.set_option("gemm_variant", variant)
q = torch.randn(nq, d, device="cuda"); q = q / q.norm(dim=1, keepdim=True)
D = torch.empty(nq, k, device="cuda"); I = torch.empty(nq, k, dtype=torch.int64, device="cuda")
for _ in range(2):
    eng.search_device(ws, q.data_ptr(), nq, k, D.data_ptr(), I.data_ptr(), stream=torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
