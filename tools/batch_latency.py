import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import minivectordb_b200 as mv
from minivectordb_b200 import synth
n, d, k = int(sys.argv[1]), int(sys.argv[2]), 10
eng = mv.FlatIPEngine(d)
eng.add_synthetic(1234, 0, n, 0, True)
rng = np.random.default_rng(0)
h = eng.mask_handle(synth.synth_mask(100, n, 0.5))
def t(fn, reps=5):
    fn(); ts = []
    for _ in range(reps):
        a = time.perf_counter(); fn(); ts.append(time.perf_counter() - a)
    return sorted(ts)[len(ts) // 2] * 1e3
for nq in (1, 8, 9, 32, 64):
    q = rng.standard_normal((nq, d)).astype(np.float32)
    print(nq, "nomask %.2f ms" % t(lambda: eng.search(q, k, normalize=True)), "handle %.2f ms" % t(lambda: eng.search(q, k, mask=h, normalize=True)), flush=True)
eng.remove_rows(np.arange(0, 1000))
q = rng.standard_normal((32, d)).astype(np.float32)
print("32 with tombstones %.2f ms" % t(lambda: eng.search(q, k, normalize=True)))
eng.add(rng.standard_normal((16, d)).astype(np.float32), normalize=True)
a = time.perf_counter(); eng.search(q, k, normalize=True); print("32 after add %.2f ms" % ((time.perf_counter() - a) * 1e3))
import threading
def many(nt):
    lat = []
    def w(i):
        qq = rng.standard_normal((1, d)).astype(np.float32)
        for _ in range(10):
            a = time.perf_counter(); eng.search(qq, k, mask=h if i % 2 else None, normalize=True); lat.append(time.perf_counter() - a)
    ts = [threading.Thread(target=w, args=(i,)) for i in range(nt)]
    a = time.perf_counter(); [x.start() for x in ts]; [x.join() for x in ts]; dt = time.perf_counter() - a
    print(nt, "threads: %.0f QPS p50 %.1f ms" % (nt * 10 / dt, np.median(lat) * 1e3), flush=True)
for nt in (8, 32, 64):
    many(nt)
