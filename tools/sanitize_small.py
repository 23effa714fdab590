"""Small end-to-end exercise of the round-2 kernels for compute-sanitizer (memcheck / racecheck / synccheck):
fp32 scan with the extraction tails (k = 1, 10, 16) and the sort tails (k = 100), the int8 shadow scan incl. its
overflow fallback, a 3-shard group on one device (fused exchange), filters and tombstones."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import minivectordb_b200 as mv
from oracle import oracle as O
n, d = 40_000, 128
x = O.synth_rows(1, 0, n, d); O.normalize_L2(x)
q = O.synth_rows(2, 0, 3, d); O.normalize_L2(q)
eng = mv.FlatIPEngine(d); eng.set_option("coalesce", 0); eng.add(x)
if os.environ.get("MVDB_HOST_PATH"): eng.set_option("host_path", int(os.environ["MVDB_HOST_PATH"]))
adm = np.random.default_rng(0).random(n) < 0.4
for shadow in (0, 1):
    eng.set_option("scan_shadow", shadow)
    for k in (1, 10, 16, 100):
        for m in (None, adm):
            D, I = eng.search(q[:1], k, mask=m)
            assert I.max() < n, (shadow, k, m is not None, I, D)
            Dr, Ir = (O.search_flat_ip(x, q[:1], k) if m is None else O.search_masked(x, m, q[:1], k))
            assert O.classify_parity(x, q[:1], I, D, Ir, Dr, admissible=m)["ok"], (shadow, k)
eng.remove_rows(np.arange(0, n, 7))
D, I = eng.search(q[:1], 10)
D3, I3 = eng.search(q, 10)           # 3 queries: multi kernel
eng.close()
engs = [mv.FlatIPEngine(d) for _ in range(3)]
for i, e in enumerate(engs):
    e.add(x[i * 10000:(i + 1) * 10000 + (10000 if i == 2 else 0)])
grp = mv.ShardGroup(engs)
for k in (1, 10, 100):
    D, S, R = grp.search(q[:2], k)
grp.close()
[e.close() for e in engs]
print("sanitize_small ok")
