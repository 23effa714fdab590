"""What does ONE coalesced pass cost at the C5 shape (10M x 768, batch of 8 / 32, k = 10)?  Quiescent, then with
a deleted prefix (tombstones), a common filter, and per-query filters (mask handles through 8 threads)."""
import json, os, sys, threading, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import minivectordb_b200 as mv
from minivectordb_b200 import synth, _native
n, d, k = 10_000_000, 768, 10
eng = mv.FlatIPEngine(d, capacity_hint=n + 100_000)
eng.add_synthetic(1234, 0, n, 0, True)
rng = np.random.default_rng(0)
q = rng.standard_normal((64, d)).astype(np.float32)
out = []
def timeit(label, fn, reps=8):
    fn(); fn()
    l0 = _native.launch_count()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    dt = (time.perf_counter() - t0) / reps
    rec = dict(case=label, ms=round(dt * 1e3, 3), launches_per_call=(_native.launch_count() - l0) / reps)
    out.append(rec); print(json.dumps(rec), flush=True)
adm = synth.synth_mask(100, n, 0.5)
for nq in (8, 32):
    timeit(f"quiescent, batch {nq}", lambda: eng.search(q[:nq], k, normalize=True))
    timeit(f"quiescent, batch {nq}, common 50% filter", lambda: eng.search(q[:nq], k, mask=adm, normalize=True))
eng.remove_rows(np.arange(0, 2_000_000))
for nq in (8, 32):
    timeit(f"2M oldest rows deleted, batch {nq}", lambda: eng.search(q[:nq], k, normalize=True))
    timeit(f"2M oldest rows deleted, batch {nq}, common filter", lambda: eng.search(q[:nq], k, mask=adm, normalize=True))
handles = [eng.mask_handle(synth.synth_mask(100 + i, n, 0.5)) for i in range(4)]
def threaded(nthreads, per):
    def run(t):
        for i in range(per):
            eng.search(q[t:t + 1], k, mask=handles[t % 4] if t % 2 else None, normalize=True)
    ts = [threading.Thread(target=run, args=(t,)) for t in range(nthreads)]
    [t.start() for t in ts]; [t.join() for t in ts]
for nt in (8, 32):
    threaded(nt, 3)
    l0 = _native.launch_count(); t0 = time.perf_counter(); threaded(nt, 20); dt = time.perf_counter() - t0
    rec = dict(case=f"{nt} threads x 20 single queries, 50% with their own mask handle (coalesced), tombstones, NO churn",
               qps=round(nt * 20 / dt, 1), ms_per_round=round(dt / 20 * 1e3, 3), launches_per_round=(_native.launch_count() - l0) / 20)
    out.append(rec); print(json.dumps(rec), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/c5_pass_probe.json", "w"), indent=1)
