"""Where do the GEMM kernels wait?  Per-CTA cycle counters (option gemm_prof) of the LAST launch of a
batched search (the final 1/2 piece), averaged over CTAs, for each kernel variant."""
import os, sys, json, ctypes
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import minivectordb_b200 as mv
from minivectordb_b200 import _native as N
out = []
cases = ((2_000_000, 1024, 4096, 100), (1_000_000, 384, 4096, 10))
for n, d, nq, k in cases:
    eng = mv.FlatIPEngine(d); eng.add_synthetic(1234, 0, n, 0, True); ws = eng.workspace()
    q = torch.randn(nq, d, device="cuda"); q = q / q.norm(dim=1, keepdim=True)
    D = torch.empty(nq, k, device="cuda"); I = torch.empty(nq, k, dtype=torch.int64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for variant, dbg in [tuple(int(x) for x in v.split(":")) for v in os.environ.get("VARIANTS", "2:0,2:1,2:2,2:3,1:0,1:1,1:3").split(",")]:
        eng.set_option("gemm_variant", variant)
        eng.set_option("gemm_debug", dbg)
        eng.set_option("gemm_prof", 0)
        for _ in range(2):
            eng.search_device(ws, q.data_ptr(), nq, k, D.data_ptr(), I.data_ptr(), stream=st)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); eng.search_device(ws, q.data_ptr(), nq, k, D.data_ptr(), I.data_ptr(), stream=st); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        eng.set_option("gemm_prof", 1)
        eng.search_device(ws, q.data_ptr(), nq, k, D.data_ptr(), I.data_ptr(), stream=st); torch.cuda.synchronize()
        buf = np.zeros((148, 8), dtype=np.uint64)
        N.check(N.lib().mvdb_debug_read_gemm_prof(eng._h, buf.ctypes.data, 148))
        b = buf.astype(np.float64)
        lead = b[b[:, 0] > 0]          # CTAs that issue MMAs (all of them, or the pair leaders)
        rec = dict(n=n, d=d, nq=nq, variant=variant, debug=dbg, ms=round(ms, 2), mma_ns=int(lead[:, 7].mean()), ghz=round(float(lead[:, 0].mean() / lead[:, 7].mean()), 3),
                   mma_total=int(lead[:, 0].mean()), mma_wait_full=int(lead[:, 1].mean()), mma_wait_tempty=int(lead[:, 2].mean()),
                   prod_total=int(b[:, 3].mean()), prod_wait_empty=int(b[:, 4].mean()),
                   epi_total=int(b[:, 5].mean()), epi_wait_tfull=int(b[:, 6].mean()),
                   mma_total_minmax=[int(lead[:, 0].min()), int(lead[:, 0].max())])
        out.append(rec); print(json.dumps(rec), flush=True)
    del ws; eng.close()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/gemm_prof_probe.json", "w"), indent=1)
