"""One batched (tensor-core) search configuration, run a few times (for ncu launch lists):
python tools/batched_one.py n d nq k [gemm_variant] [batch_mode] [iters]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import minivectordb_b200 as mv
n, d, nq, k = (int(a) for a in sys.argv[1:5])
variant = int(sys.argv[5]) if len(sys.argv) > 5 else 2
mode = int(sys.argv[6]) if len(sys.argv) > 6 else 1
iters = int(sys.argv[7]) if len(sys.argv) > 7 else 2
eng = mv.FlatIPEngine(d); eng.add_synthetic(1234, 0, n, 0, True)
eng.set_option("gemm_variant", variant); eng.set_option("batch_mode", mode)
ws = eng.workspace()
q = torch.randn(nq, d, device="cuda"); q = q / q.norm(dim=1, keepdim=True)
D = torch.empty(nq, k, device="cuda"); I = torch.empty(nq, k, dtype=torch.int64, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for _ in range(iters):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); eng.search_device(ws, q.data_ptr(), nq, k, D.data_ptr(), I.data_ptr(), stream=st); e1.record()
    torch.cuda.synchronize()
    print("search ms", round(e0.elapsed_time(e1), 3), flush=True)
