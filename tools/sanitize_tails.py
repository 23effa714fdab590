"""Second compute-sanitizer exercise (memcheck / synccheck / initcheck): the block-wide selection tails of the
survivor-list scan (32 < k <= 128) and of the int8 shadow scan, incl. list overflow (thousands of exact ties ->
classic-scan fallback), the fused exchange behind those tails (3-shard group on one device), the histogram select for k > 128, and the host path
(inputs pulled by a grid the scan depends on programmatically, pinned caller filter, results in pinned memory)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import minivectordb_b200 as mv
from oracle import oracle as O
n, d = 60_000, 64
x = O.synth_rows(11, 0, n, d); O.normalize_L2(x)
q = O.synth_rows(12, 0, 2, d); O.normalize_L2(q)
adm = np.random.default_rng(1).random(n) < 0.3
pinned = torch.from_numpy(mv.pack_mask(adm)).pin_memory()
eng = mv.FlatIPEngine(d); eng.set_option("coalesce", 0); eng.add(x)
for shadow in (0, 1):
    eng.set_option("scan_shadow", shadow)
    for k in (33, 64, 100, 128):
        for m in (None, adm, pinned.numpy()):
            D, I = eng.search(q[:1], k, mask=m, mask_rows=n if (m is not None and m.dtype != np.bool_) else None)
            assert I.max() < n, (shadow, k)
            a = None if m is None else adm
            Dr, Ir = (O.search_flat_ip(x, q[:1], k) if a is None else O.search_masked(x, a, q[:1], k))
            assert O.classify_parity(x, q[:1], I, D, Ir, Dr, admissible=a)["ok"], (shadow, k)
# list overflow: 9000 copies of one row tie at the top
eng.add(np.repeat(x[7:8], 9000, axis=0))
for shadow in (0, 1):
    eng.set_option("scan_shadow", shadow)
    for k in (10, 100):
        D, I = eng.search(x[7:8], k)
        assert I.max() < n + 9000 and np.all(D > 0.999), (shadow, k)
# k > 128 on the host path: histogram select against the radix select, incl. its overflow (9000 ties is below its cap,
# 20000 more are not)
for extra in (0, 20000):
    if extra: eng.add(np.repeat(x[7:8], extra, axis=0))
    for qq in (q[:1], x[7:8]):
        for k in (200, 1000):
            eng.set_option("large_k_fast", 0); Dr, Ir = eng.search(qq, k)
            eng.set_option("large_k_fast", 1); Dg, Ig = eng.search(qq, k)
            assert np.array_equal(Ir, Ig) and np.array_equal(Dr, Dg), (extra, k)
eng.close()
# fused exchange behind the selection tails
for shadow in (0, 1):
    engs = [mv.FlatIPEngine(d) for _ in range(3)]
    for i, e in enumerate(engs):
        e.set_option("scan_shadow", shadow)
        e.add(x[i * 20000:(i + 1) * 20000])
    grp = mv.ShardGroup(engs)
    for k in (64, 100):
        D, S, R = grp.search(q[:1], k)
        Dr, Ir = O.search_flat_ip(x, q[:1], k)
        assert O.classify_parity(x, q[:1], S * 20000 + R, D, Ir, Dr)["ok"], (shadow, k)
    grp.close()
    [e.close() for e in engs]
print("sanitize_tails ok")
