#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 flat inner-product engine.

  python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload c4|c2|c1|c5|c3]

Default workload, at every N (BASELINE.json configs[3], the north-star target):

  C4 shard -- 12.5M x 512 fp32 rows per GPU, one unfiltered query, k = 10.
  N = 8 is exactly "ShardedVectorDatabase 100M x 512 row-sharded across 8 x B200"; N = 1 is the
  largest single-GPU top-10 configuration that BASELINE.json names (25.6 GB).  One "step" = one
  query = one pass of the hot path over every resident row.

N > 1 is launched by torchrun (one rank per GPU): every rank holds its own shard, scans it, and the
per-GPU top-k are exchanged and merged inside the scan kernel over NVLink peer stores (fused
exchange; `--exchange nccl` = all-gather + merge kernel).  Weak scaling: rows per GPU are fixed.
`value` counts shard scans per second over all ranks (= N x global QPS); the global QPS and the
p50 latency are reported next to it.

Rank 0 prints ONE JSON line.  Besides the contract's keys it carries
  roofline      algorithmic bytes / isolated launch time vs MEASURED_PEAKS.json, per GPU and aggregate
  parity        computed OUTSIDE the timed region: every rank's final (D, I) identical, and equal (the
                parity rule of oracle.classify_parity) to the CPU oracle streamed over 1M-row chunks of
                every shard for 4 of the timed queries
  cpu_baseline  the oracle on this box's host cores, bounded sample (N = 1 only)
  secondary     BASELINE.json configs[1] (C2: 1M x 384, ~50 % filter bitmask) measured the same way
See DESIGN.md "Measurement" for every field.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED_DB, SEED_Q, SEED_META = 1234, 4321, 99
TOPK = 10
PARITY_QUERIES = 4
ORACLE_CHUNK = 1 << 20

WORKLOADS = {
    "c4": dict(rows=12_500_000, dim=512, filt=False,
               label="C4 shard: 12.5M x 512 fp32 per GPU, nq=1, k=10, unfiltered "
                     "(x8 GPUs = BASELINE.json configs[3], 100M x 512 row-sharded)"),
    "c2": dict(rows=1_000_000, dim=384, filt=True,
               label="C2: 1M x 384 fp32 per GPU, nq=1, k=10, ~50% metadata-filter bitmask (BASELINE.json configs[1])"),
    "c1": dict(rows=100_000, dim=512, filt=False, flush=True,
               label="C1: 100k x 512 fp32, nq=1, k=10, unfiltered (BASELINE.json configs[0]); L2 flushed between queries"),
    "c5": dict(rows=10_000_000, dim=768, filt=True,
               label="C5 shape, quiesced: 10M x 768 fp32, nq=1, k=10, ~50% filter bitmask (BASELINE.json configs[4] "
                     "without the churn threads; the churn run is tools/churn_probe.py)"),
}
METRIC = "QPS flat-IP top-10, single query (shard scans/s over all GPUs)"


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            j = json.load(f)
        return float(j["hbm_gbs"]), float(j.get("bf16_tflops", 1668.4)), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, 1590.0, "fallback (B200_PROFILING.md: 6.65 TB/s, 1.59 PFLOP/s)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                pw.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": float(max(pw))}


def make_config(wl_name, world, exchange=None, no_pdl=False):
    """The `config` object -- identical for the B200 arm and the reference arm of one (workload, N)."""
    wl = WORKLOADS[wl_name]
    n, d = wl["rows"], wl["dim"]
    return {"workload": wl["label"], "rows_per_gpu": n, "rows_total": n * world, "dim": d, "k": TOPK, "nq": 1,
            "filter_keep": 0.5 if wl["filt"] else None,
            "l2": (f"working set {n * d * 4 / 1e9:.2f} GB per GPU vs 126 MB L2, distinct query every step: "
                   + ("L2 flushed (256 MB write) before every isolated launch" if wl.get("flush") else "no flush needed")),
            "unit_note": "value = shard scans/s over all ranks = n_gpus x global QPS (every rank scans its own "
                         f"{n} x {d} shard per query); qps_global = queries/s over the whole {n * world}-row database"}


# ---------------------------------------------------------------------------
# CPU arm: the oracle (C restatement of faiss's flat-IP scan) on the host cores
# ---------------------------------------------------------------------------
def cpu_scan_qps(wl_name, rows_total, budget_s, reps, warm, threads, max_rows=None):
    """What the reference does per query on its faiss-cpu path (ref vector_database.py:497, 508-514): one
    single-threaded scan per query (faiss parallelises over queries only), `threads` queries at a time on
    all host cores; with a filter, first gather the admissible rows into a temporary index.  The matrix is
    the workload's own (same generator, same seeds).  When `rows_total` rows do not fit the host (or the
    time budget), a prefix is scanned and the rate is scaled linearly -- said so in `sample`."""
    from oracle import oracle as O
    from minivectordb_b200 import synth
    import concurrent.futures as cf
    wl = WORKLOADS[wl_name]
    d = wl["dim"]
    try:
        import psutil
        avail = psutil.virtual_memory().available
    except Exception:
        avail = 32 << 30
    fit = int(avail * 0.6 / (d * 4 * (1.0 + (0.5 * threads if wl["filt"] else 0.0))))   # the gather branch copies ~half the rows, per thread
    rows = min(rows_total, fit, max_rows or rows_total)
    if os.environ.get("MVDB_BENCH_CPU_ROWS"):   # tests: keep the CPU arm small
        rows = min(rows, int(os.environ["MVDB_BENCH_CPU_ROWS"]))
    x = np.empty((rows, d), dtype=np.float32)
    for r0 in range(0, rows, ORACLE_CHUNK):
        m = min(ORACLE_CHUNK, rows - r0)
        O.synth_rows(SEED_DB, r0, m, d, out=x[r0:r0 + m])
    O.normalize_L2(x)
    sel = np.flatnonzero(synth.synth_mask(SEED_META, rows, 0.5)).astype(np.int64) if wl["filt"] else None
    q = O.synth_rows(SEED_Q, 0, threads * (reps + max(warm, 1)), d)
    O.normalize_L2(q)

    def step(s, nrows):
        def one(i):
            if sel is None:
                return O.search_flat_ip(x[:nrows], q[i:i + 1], TOPK, nthreads=1)
            return O.search_gathered(x, sel[:np.searchsorted(sel, nrows)], q[i:i + 1], TOPK, nthreads=1)
        with cf.ThreadPoolExecutor(max_workers=threads) as ex:
            list(ex.map(one, range(s * threads, (s + 1) * threads)))

    # a calibration step (it is also the first warm-up) decides whether every step can afford all `rows`
    t0 = time.perf_counter()
    step(0, rows)
    t_step = time.perf_counter() - t0
    use = rows
    if t_step * (reps + warm) > budget_s:
        use = min(rows, max(ORACLE_CHUNK // 4, int(rows * budget_s / (t_step * (reps + warm)))))
    for w in range(1, warm):
        step(w, use)
    t0 = time.perf_counter()
    for s in range(reps):
        step(max(warm, 1) + s, use)
    dt = time.perf_counter() - t0
    qps_sample = reps * threads / dt
    qps = qps_sample * use / rows_total
    sample = (f"{reps} steps x {threads} concurrent single-thread queries "
              f"({'gather admissible rows + ' if wl['filt'] else ''}scan), each over {use} of the {rows_total} rows"
              + ("" if use == rows_total else f"; rate scaled linearly by {use}/{rows_total} (sample: host RAM / time budget)"))
    return qps, dt / reps, sample, dict(rows_scanned=use, host_rows_resident=rows, qps_on_sample=qps_sample)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = args.gpus
    threads = os.cpu_count() or 1
    wl = WORKLOADS[args.workload]
    # the host keeps at most ONE shard's rows resident; a larger database (N > 1) is scaled from it
    qps, s_per_step, sample, extra = cpu_scan_qps(args.workload, wl["rows"] * world, budget_s=150.0, reps=args.steps,
                                                  warm=args.warmup, threads=threads, max_rows=wl["rows"])
    line = {
        "impl": "reference", "metric": METRIC, "value": qps * world, "unit": "queries/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": s_per_step * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": make_config(args.workload, world),
        "qps_global": qps,
        "cpu_baseline": {"value": qps * world, "unit": "queries/s", "cores": threads, "kind": "port", "sample": sample,
                         "note": "oracle port of faiss IndexFlatIP (faiss-cpu is not installable here: DESIGN.md section 5); "
                                 "host cores only", **extra},
        "e2e": {"value": qps * world, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def cpu_baseline_sample(wl_name):
    """Oracle on the box's host cores, bounded sample of the same workload (10-30 s of CPU work)."""
    threads = os.cpu_count() or 1
    wl = WORKLOADS[wl_name]
    qps, _, sample, extra = cpu_scan_qps(wl_name, wl["rows"], budget_s=12.0, reps=2, warm=1, threads=threads,
                                         max_rows=2_000_000)
    out = {"value": qps, "unit": "queries/s", "cores": threads, "kind": "port", "sample": sample, **extra}
    try:   # the unmodified reference classes (Python filter + gather + rebuild), measured in the build container
        with open(os.path.join(ROOT, "profiles", "r02_reference_python_path.json")) as f:
            out["reference_python_path"] = json.load(f)
    except Exception:
        pass
    return out


# ---------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------
def measure_single_query(args, wl_name, rank, world, local, with_parity=True, clocks=None, shadow=False):
    """One workload on this rank's GPU: device-resident leg (value, roofline), end-to-end leg through
    the C ABI with host buffers, parity against the streamed oracle.  Collective under torchrun."""
    import torch
    import torch.distributed as dist
    import minivectordb_b200 as mv
    from minivectordb_b200 import _native, synth
    from minivectordb_b200.distributed import RowShardedIndex

    wl = WORKLOADS[wl_name]
    n, d, k = (args.rows or wl["rows"]), (args.dim or wl["dim"]), TOPK
    filt = wl["filt"] and not args.no_filter
    K, W = args.steps, args.warmup

    index = RowShardedIndex(d, device=local, exchange=args.exchange)   # one row shard per rank, resident in HBM
    index.add(synthetic=(SEED_DB, rank * n, n, 0), normalize=True)
    eng = index.engine
    ld = eng.device_view()[1]
    if shadow:
        # opt-in int8 shadow mode: int8 rows + exact fp32 re-scoring of a rigorous candidate superset; the
        # answers must be (and are checked to be) bit-identical to the fp32 scan's
        eng.set_option("scan_shadow", 1)
    adm = packed = words = mask_dev = None
    if filt:
        adm = synth.synth_mask(SEED_META + rank, n, 0.5)
        packed = mv.pack_mask(adm)
        words = np.zeros((n + 31) // 32 * 4, dtype=np.uint8)
        words[:packed.size] = packed
        mask_dev = torch.from_numpy(words.view(np.int32)).cuda()
    q_host = synth.synth_rows(SEED_Q, 0, K + W, d)
    q_host /= np.linalg.norm(q_host, axis=1, keepdims=True)
    q_host = np.ascontiguousarray(q_host, dtype=np.float32)
    q_dev = torch.from_numpy(q_host).cuda()
    stream = torch.cuda.current_stream()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda") if wl.get("flush") else None

    def step(i):
        # scan this rank's shard; N > 1: the scan's last CTA exchanges the k best over NVLink and merges
        if filt:
            return index.search_device(q_dev[i:i + 1], k, mask_dev, n)
        return index.search_device(q_dev[i:i + 1], k)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident leg: `value` + roofline ----------------------------
    # Timed region: EXACTLY K searches enqueued back to back on one stream, bracketed by one event
    # pair.  With option "pdl" (programmatic dependent launch) the scan of search i+1 fills the SMs
    # that search i has left, so the serial tail of a search (last-CTA merge, cross-GPU exchange)
    # overlaps the next scan; results are unchanged.  A second pass with an event after every
    # search (which serialises the launches) gives the per-launch duration for the roofline and p50.
    eng.set_option("pdl", 0 if args.no_pdl else 1)
    for i in range(W):
        step(i)
    barrier()
    launches0 = _native.launch_count()
    e_begin, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e_begin.record(stream)
    for i in range(K):
        step(W + i)
    e_end.record(stream)
    barrier()
    launches = _native.launch_count() - launches0
    total_s = max_over_ranks(e_begin.elapsed_time(e_end) * 1e-3)
    # per-launch durations (isolated launches: the event between two searches serialises them)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    barrier()
    for i in range(K):
        if flush is not None:
            flush.zero_()           # evict the matrix from L2: every query finds it cold (config < L2-resident sizes)
        ev[i][0].record(stream)
        step(W + i)
        ev[i][1].record(stream)
    barrier()
    per_step = np.array([a.elapsed_time(b) * 1e-3 for a, b in ev])
    iso_mean = max_over_ranks(float(np.mean(per_step)))
    iso_p50 = max_over_ranks(float(np.median(per_step)))
    eng.set_option("pdl", 0)
    index.check_exchange()
    # the device leg's answers for the parity check (same queries as the first timed steps)
    dev_res = []
    for i in range(min(PARITY_QUERIES, K)):
        D, I = step(W + i)
        torch.cuda.synchronize()
        dev_res.append((D.cpu().numpy().copy(), I.cpu().numpy().copy()))

    # ---- end-to-end leg: host buffers through the C ABI ------------------------
    # every step copies the query (d*4 B) and, with a filter, the packed bitmask (n/8 B) H2D and the
    # (D, I) result D2H inside the timed region.
    e2e_lat, e2e_res = [], []
    # the step's inputs lie in pinned host memory (what the contract asks for): the library then pulls the filter
    # from where it lies instead of staging it through its own pinned buffer first
    keep_pinned = []
    if world == 1:
        tq = torch.from_numpy(q_host).pin_memory()
        keep_pinned.append(tq)
        q_e2e = tq.numpy()
        if filt:
            tm = torch.from_numpy(packed).pin_memory()
            keep_pinned.append(tm)
            packed = tm.numpy()
    else:
        q_e2e = q_host

    def e2e_step(i):
        if world == 1:
            if filt:
                return eng.search(q_e2e[i:i + 1], k, mask=packed, mask_rows=n)
            return eng.search(q_e2e[i:i + 1], k)
        # one pinned H2D ([filter words | query]), scan + fused exchange + merge, one D2H ([labels | distances])
        if filt:
            return index.search_packed(q_host[i:i + 1], k, words, n)
        return index.search_packed(q_host[i:i + 1], k)

    for i in range(W):
        e2e_step(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(K):
        a = time.perf_counter()
        r = e2e_step(W + i)
        e2e_lat.append(time.perf_counter() - a)
        if i < PARITY_QUERIES:
            e2e_res.append(r)
    barrier()
    e2e_total = max_over_ranks(time.perf_counter() - t0)
    clk = clocks.stop() if clocks is not None else None   # clocks cover the timed legs only, not the CPU oracle below

    # ---- parity (outside every timed region) -----------------------------------
    parity = None
    if with_parity and not args.no_parity:
        parity = parity_check(rank, world, n, d, k, adm, q_host[W:W + len(dev_res)], dev_res, e2e_res)

    res = dict(n=n, d=d, ld=ld, k=k, filt=filt, K=K, W=W, total_s=total_s, iso_mean=iso_mean, iso_p50=iso_p50,
               launches=int(launches), e2e_total=e2e_total, e2e_p50=float(np.median(e2e_lat)), parity=parity,
               exchange=index.exchange, clocks=clk, dev_res=dev_res,
               kernel="scan_i8_kernel" if shadow else f"scan_q1_kernel<{(ld // 4 + 31) // 32},tma>",
               # N*d*4 (+ ceil(N/8) with a filter mask); shadow mode streams N*(d16 + 16) bytes of int8 records
               alg_bytes=(n * ((d + 15) // 16 * 16 + 16) if shadow else n * ld * 4) + ((n + 7) // 8 if filt else 0),
               h2d=d * 4 + ((n + 7) // 8 if filt else 0), d2h=k * 12, flush=flush is not None)
    index.close()
    del q_dev, mask_dev, flush
    torch.cuda.empty_cache()
    return res


def parity_check(rank, world, n, d, k, adm, q, dev_res, e2e_res):
    """(1) every rank ended with the same (D, I), on the device leg and on the e2e leg, and the two legs
    agree; (2) that answer equals -- by the parity rule (ids position-wise, mismatch excused only for
    exact / fp32-near ties judged in float64, distances <= 1e-5 relative) -- the CPU oracle streamed over
    1M-row chunks of EVERY shard, merged on the host."""
    import torch.distributed as dist
    from oracle import oracle as O   # test infrastructure: the checker, never the thing measured
    nq = len(dev_res)
    D_dev = np.concatenate([r[0] for r in dev_res])
    I_dev = np.concatenate([r[1] for r in dev_res])
    D_e2e = np.concatenate([r[0] for r in e2e_res]) if e2e_res else D_dev
    I_e2e = np.concatenate([r[1] for r in e2e_res]) if e2e_res else I_dev
    t0 = time.perf_counter()
    threads = max(1, (os.cpu_count() or 1) // world)
    buf = np.empty((min(ORACLE_CHUNK, n), d), dtype=np.float32)

    def chunk(r0, m):
        x = O.synth_rows(SEED_DB, rank * n + r0, m, d, out=buf)
        O.normalize_L2(x)
        return x

    D_loc, I_loc = O.search_streamed(chunk, n, q, k, chunk_rows=ORACLE_CHUNK, row_offset=rank * n, admissible=adm,
                                     nthreads=min(threads, nq))
    mine = dict(D_dev=D_dev, I_dev=I_dev, D_e2e=D_e2e, I_e2e=I_e2e, D_orc=D_loc, I_orc=I_loc)
    if world > 1:
        allr = [None] * world
        dist.all_gather_object(allr, mine)
    else:
        allr = [mine]
    if rank != 0:
        return None
    same_dev = all(np.array_equal(r["I_dev"], allr[0]["I_dev"]) and np.array_equal(r["D_dev"], allr[0]["D_dev"]) for r in allr)
    same_e2e = all(np.array_equal(r["I_e2e"], allr[0]["I_e2e"]) and np.array_equal(r["D_e2e"], allr[0]["D_e2e"]) for r in allr)
    legs_agree = np.array_equal(I_dev, I_e2e) and np.array_equal(D_dev, D_e2e)
    D_ref, I_ref = O.merge_topk_lists([(r["D_orc"], r["I_orc"]) for r in allr], k)

    def fetch(label):
        x = O.synth_rows(SEED_DB, int(label), 1, d)
        O.normalize_L2(x)
        return x[0]

    rep = O.classify_parity_lazy(fetch, d, q, I_dev, D_dev, I_ref, D_ref, rel_tol=1e-5)
    ok = bool(rep["ok"] and same_dev and same_e2e and legs_agree)
    return {"ok": ok, "queries": nq, "k": k, "rows_checked": n * world, "oracle": f"C restatement streamed over {ORACLE_CHUNK}-row "
            f"chunks of every shard ({world} shard(s)), per-chunk top-k merged on the host", "ranks_identical_device_leg": bool(same_dev),
            "ranks_identical_e2e_leg": bool(same_e2e), "device_and_e2e_legs_identical": bool(legs_agree),
            "ids_equal": rep["id_equal"], "positions": rep["positions"], "exact_ties": rep["exact_tie"],
            "near_ties": rep["near_tie"], "real_errors": rep["real_error"], "max_rel_err": rep["max_rel_err"],
            "rel_tol": 1e-5, "oracle_seconds": round(time.perf_counter() - t0, 2)}


def roofline_of(res, world, peak, peak_src, wl_name):
    ach = res["alg_bytes"] / res["iso_mean"] / 1e9
    ach_pipe = res["alg_bytes"] / (res["total_s"] / res["K"]) / 1e9
    out = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
           "kernel": res["kernel"], "alg_bytes_per_launch": res["alg_bytes"], "avg_launch_us": res["iso_mean"] * 1e6,
           "peak_source": peak_src,
           "note": "per GPU; achieved/frac use the ISOLATED per-launch duration (an event pair around every search, max over "
                   "ranks of the mean; N > 1: includes the in-kernel NVLink exchange); `value` is the back-to-back rate of "
                   "the timed region, where consecutive launches overlap (pipelined_*). The denominator is a read+write "
                   "copy peak, a read-only stream can exceed it.",
           "pipelined_achieved": ach_pipe, "pipelined_frac": ach_pipe / peak,
           "aggregate": {"n_gpus": world, "achieved": ach * world, "peak": peak * world, "frac": ach / peak,
                         "pipelined_achieved": ach_pipe * world, "pipelined_frac": ach_pipe / peak}}
    # DRAM bytes per launch from the committed `ncu --set full` capture of this workload -- only while the
    # kernel sources are the ones that were profiled
    try:
        import hashlib
        prof = json.load(open(os.path.join(ROOT, "profiles", "scan_traffic.json"))).get(wl_name)
        h = hashlib.sha256()
        for f in ("scan.cuh", "select.cuh", "device_utils.cuh"):
            h.update(open(os.path.join(ROOT, "minivectordb_b200", "csrc", f), "rb").read())
        if prof and prof.get("rows") == res["n"] and prof.get("dim") == res["d"]:
            out["traffic"] = prof.get("dram_bytes_per_launch")
            out["traffic_source"] = prof.get("source")
            if prof.get("kernel_sources_sha256") != h.hexdigest()[:16]:
                out["traffic_note"] = "captured with an earlier revision of the scan kernel sources"
    except Exception:
        pass
    return out


def run_b200(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world == 1 and args.gpus > 1:
        raise SystemExit("--gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    peak, _, peak_src = measured_peaks()
    custom = bool(args.rows or args.dim or args.no_filter)

    clocks = None
    if rank == 0:
        clocks = ClockSampler(local)
        clocks.start()
    res = measure_single_query(args, args.workload, rank, world, local, clocks=clocks)
    clk = res["clocks"]
    sec = shd = None
    if args.workload == "c4" and not custom and not args.no_secondary:
        sec = measure_single_query(args, "c2", rank, world, local)
        shd = measure_single_query(args, "c4", rank, world, local, shadow=True)

    if rank == 0:
        K, W = res["K"], res["W"]
        qps_global = K / res["total_s"]
        cfg = make_config(args.workload, world)
        if custom:
            cfg.update(workload=f"custom: {res['n']} x {res['d']} fp32 per GPU, nq=1, k={TOPK}, "
                       + ("~50% filter bitmask" if res["filt"] else "no filter"),
                       rows_per_gpu=res["n"], rows_total=res["n"] * world, dim=res["d"],
                       filter_keep=0.5 if res["filt"] else None)
        line = {
            "metric": METRIC, "value": qps_global * world, "unit": "queries/s", "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": res["total_s"] / K * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfg,
            "run": {"launch": "plain stream order" if args.no_pdl else
                              "programmatic dependent launch: back-to-back searches overlap one search's merge tail / exchange with the next scan",
                    "parallelism": f"row-shard x{world}" + (f", exchange={res['exchange']}" if world > 1 else "")},
            "qps_global": qps_global,
            "p50_latency_us": res["iso_p50"] * 1e6,
            "e2e": {"value": K / res["e2e_total"] * world, "unit": "queries/s",
                    "h2d_bytes_per_step": res["h2d"], "d2h_bytes_per_step": res["d2h"],
                    "p50_latency_us": res["e2e_p50"] * 1e6, "qps_global": K / res["e2e_total"],
                    "api": ("mvdb_index_search (C ABI, host buffers; every step the query" + ("+mask" if res["filt"] else "") + " cross PCIe from pinned host memory -- pulled by a grid the scan is launched behind -- and the results are written by the scan into pinned host memory and copied to the caller's arrays)"
                            if world == 1 else
                            "RowShardedIndex.search_packed: one pinned H2D -> scan + fused NVLink exchange + merge -> one D2H")},
            "gpu_launches": res["launches"],
            "clocks": clk,
            "roofline": roofline_of(res, world, peak, peak_src, args.workload),
            "parity": res["parity"],
        }
        if sec is not None:
            sq = sec["K"] / sec["total_s"]
            line["secondary"] = {
                "config": make_config("c2", world), "value": sq * world, "qps_global": sq,
                "ms_per_step": sec["total_s"] / sec["K"] * 1e3, "p50_latency_us": sec["iso_p50"] * 1e6,
                "e2e": {"value": sec["K"] / sec["e2e_total"] * world, "p50_latency_us": sec["e2e_p50"] * 1e6,
                        "h2d_bytes_per_step": sec["h2d"], "d2h_bytes_per_step": sec["d2h"]},
                "roofline": roofline_of(sec, world, peak, peak_src, "c2"), "parity": sec["parity"]}
        if shd is not None:
            sq = shd["K"] / shd["total_s"]
            same = all(np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) for a, b in zip(res["dev_res"], shd["dev_res"]))
            rf = roofline_of(shd, world, peak, peak_src, "c4_shadow")
            rf["note"] = ("int8 shadow mode: algorithmic bytes = N x (d rounded to 16 + 16) bytes of int8 records per launch (the fp32 "
                          "re-scoring of the few hundred candidates is not counted); same timing rules as the main roofline")
            line["shadow_mode"] = {
                "what": "OPT-IN mode (option scan_shadow = 1), not the headline: the query streams an int8 shadow of the matrix "
                        "(dp4a, two-plane int8 query), keeps every row a rigorous error bound cannot rule out, re-scores those "
                        "from the fp32 rows with the fp32 scan's summation order -- ids and distances bit-identical to the fp32 scan",
                "config": make_config("c4", world), "value": sq * world, "qps_global": sq,
                "ms_per_step": shd["total_s"] / shd["K"] * 1e3, "p50_latency_us": shd["iso_p50"] * 1e6,
                "e2e": {"value": shd["K"] / shd["e2e_total"] * world, "p50_latency_us": shd["e2e_p50"] * 1e6,
                        "h2d_bytes_per_step": shd["h2d"], "d2h_bytes_per_step": shd["d2h"]},
                "roofline": rf, "parity": shd["parity"], "identical_to_fp32_scan": bool(same),
                "speedup_vs_fp32_scan": (res["total_s"] / res["K"]) / (shd["total_s"] / shd["K"])}
        if world == 1 and not args.no_cpu_baseline and not custom:
            line["cpu_baseline"] = cpu_baseline_sample(args.workload)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ---------------------------------------------------------------------------
# C3: large query batches on the tensor cores (builder-run; same contract line)
# ---------------------------------------------------------------------------
def run_c3(args):
    import torch
    from minivectordb_b200 import FlatIPEngine, _native, synth
    torch.cuda.set_device(0)
    n, d, nq, k = args.rows or 10_000_000, args.dim or 1024, args.nq or 4096, 100
    mode = {"exact": 1, "bf16": 2, "tf32": 3}[args.mode]
    _, tf_peak, peak_src = measured_peaks()
    eng = FlatIPEngine(d)
    eng.add_synthetic(SEED_DB, 0, n, 0, True)
    eng.set_option("batch_mode", mode)
    eng.set_option("batch_cost_model", 0)
    ws = eng.workspace()
    K, W = args.steps, args.warmup
    q_host = synth.synth_rows(SEED_Q, 0, nq, d)
    q_host /= np.linalg.norm(q_host, axis=1, keepdims=True)
    q_host = np.ascontiguousarray(q_host, dtype=np.float32)
    q_dev = torch.from_numpy(q_host).cuda()
    D = torch.empty((nq, k), dtype=torch.float32, device="cuda")
    I = torch.empty((nq, k), dtype=torch.int64, device="cuda")
    st = torch.cuda.current_stream()

    def step():
        eng.search_device(ws, q_dev.data_ptr(), nq, k, D.data_ptr(), I.data_ptr(), stream=st.cuda_stream)

    for _ in range(W):
        step()
    torch.cuda.synchronize()
    clocks = ClockSampler(0)
    clocks.start()
    l0 = _native.launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
    ev[0].record(st)
    for i in range(K):
        step()
        ev[i + 1].record(st)
    torch.cuda.synchronize()
    launches = _native.launch_count() - l0
    per = np.array([ev[i].elapsed_time(ev[i + 1]) * 1e-3 for i in range(K)])
    t0 = time.perf_counter()
    for _ in range(max(1, K // 2)):
        De, Ie = eng.search(q_host, k)
    e2e_s = (time.perf_counter() - t0) / max(1, K // 2)
    clk = clocks.stop()
    flops = 2.0 * nq * n * d
    ach = flops / float(np.mean(per)) / 1e12
    # parity / recall on a sample of the queries: the fp32 scan of the same engine
    eng.set_option("batch_mode", 0)
    sel = np.arange(0, nq, max(1, nq // 16))[:16]
    Ds, Is = eng.search(q_host[sel], k)
    Ih = I.cpu().numpy()[sel]
    Dh = D.cpu().numpy()[sel]
    recall = float(np.mean([len(set(Ih[i]) & set(Is[i])) / k for i in range(len(sel))]))
    line = {"metric": "QPS flat-IP top-100, query batch on tensor cores", "value": nq / float(np.mean(per)), "unit": "queries/s",
            "n_gpus": 1, "steps": K, "warmup": W, "ms_per_step": float(np.mean(per)) * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": {"exact": "bf16 select + f32 re-score", "bf16": "bf16", "tf32": "tf32"}[args.mode],
            "data": "synthetic",
            "config": {"workload": f"C3: {n} x {d} fp32, nq={nq}, k={k}, mode={args.mode} (BASELINE.json configs[2])",
                       "rows_per_gpu": n, "rows_total": n, "dim": d, "k": k, "nq": nq},
            "e2e": {"value": nq / e2e_s, "unit": "queries/s", "h2d_bytes_per_step": nq * d * 4, "d2h_bytes_per_step": nq * k * 12},
            "gpu_launches": int(launches), "clocks": clk,
            "roofline": {"bound": "tensor", "achieved": ach, "peak": tf_peak, "unit": "TFLOP/s", "frac": ach / tf_peak, "traffic": None,
                         "peak_source": peak_src, "flops_per_step": flops, "kernel": "gemm_topk_kernel_mc<2> (whole search incl. threshold refreshes and re-scoring)"},
            "parity": {"sample_queries": int(len(sel)), "vs": "fp32 scan of the same engine",
                       "ids_identical": bool(np.array_equal(Ih, Is)), "distances_identical": bool(np.array_equal(Dh, Ds)),
                       "recall_at_k": recall}}
    print(json.dumps(line), flush=True)
    ws.close()
    eng.close()


def _quiet_stdout():
    """Everything that libraries print on fd 1 (e.g. NCCL's version banner) goes to
    stderr; the returned file object writes to the real stdout so that rank 0 emits
    exactly ONE JSON line there."""
    real = os.fdopen(os.dup(1), "w")
    sys.stdout.flush()
    os.dup2(2, 1)
    return real


def main():
    global print
    real_stdout = _quiet_stdout()
    _print = print

    def print(*a, **kw):  # noqa: A001 - the JSON line goes to the real stdout
        kw.setdefault("file", real_stdout)
        _print(*a, **kw)
        real_stdout.flush()

    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS) + ["c3"])
    ap.add_argument("--mode", default="exact", choices=["exact", "bf16", "tf32"], help="c3 only")
    ap.add_argument("--nq", type=int, default=0, help="c3 only: queries per batch")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU leg (profiling runs under ncu)")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle parity check (profiling runs under ncu)")
    ap.add_argument("--no-secondary", action="store_true", help="skip the secondary C2 measurement")
    ap.add_argument("--rows", type=int, default=0, help="rows per GPU (custom shape)")
    ap.add_argument("--dim", type=int, default=0)
    ap.add_argument("--no-filter", action="store_true")
    ap.add_argument("--exchange", default="auto", choices=["auto", "fused", "nccl"])
    ap.add_argument("--no-pdl", action="store_true", help="plain stream-ordered launches in the timed region (A/B)")
    args = ap.parse_args()
    if args.impl == "reference":
        if args.workload == "c3":
            print(json.dumps({"impl": "reference", "unavailable": "the reference arm covers the single-query workloads"}))
            return
        # defaults keep the CPU run in minutes: a step is `cores` full scans
        args.steps = 8 if args.steps is None else args.steps
        args.warmup = 1 if args.warmup is None else args.warmup
        run_reference(args)
        return
    if args.workload == "c3":
        args.steps = 5 if args.steps is None else args.steps
        args.warmup = max(3, 3 if args.warmup is None else args.warmup)
        run_c3(args)
        return
    args.steps = 100 if args.steps is None else args.steps
    args.warmup = max(3, 5 if args.warmup is None else args.warmup)
    run_b200(args)


if __name__ == "__main__":
    main()
