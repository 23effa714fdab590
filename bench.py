#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 flat inner-product engine.

Workload (BASELINE.json configs[1], the config the metric is quoted on):
1M x 384 fp32 (e5-small dim), single query, k=10, ~50 % metadata filter as a
bitmask.  One "step" = one query = one pass of the hot path over the matrix.

  python bench.py --gpus N --steps K --warmup W [--impl reference]

N>1 is launched by torchrun (one rank per GPU): every rank holds its own
1M x 384 row shard (global DB = N x 1M rows), scans it, the per-GPU top-k are
all-gathered over NCCL and merged on every rank (weak scaling: rows per GPU
fixed).  `value` counts shard scans per second over all ranks (N x global
QPS) so that it is the whole-job throughput of the scan; the global QPS and
p50 latency are reported next to it.

Rank 0 prints ONE JSON line.  See DESIGN.md "Measurement" for every field.
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

N_ROWS, DIM, TOPK = 1_000_000, 384, 10
SEED_DB, SEED_Q, SEED_META = 1234, 4321, 99
WORKLOAD = "C2: 1M x 384 fp32, nq=1, k=10, ~50% metadata-filter bitmask (BASELINE.json configs[1])"
METRIC = "QPS flat-IP top-10 (single query, filtered), 1M x 384 per GPU"


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


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
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------
# CPU arm: the oracle (C restatement of faiss's flat-IP scan) on the host cores
# ---------------------------------------------------------------------------
def cpu_reference_setup(n, d):
    from oracle import oracle as O  # test infrastructure, allowed here as the CPU baseline only
    x = O.synth_rows(SEED_DB, 0, n, d)
    O.normalize_L2(x)
    from minivectordb_b200 import synth
    adm = synth.synth_mask(SEED_META, n, 0.5)
    rows = np.flatnonzero(adm).astype(np.int64)
    return O, x, rows


def cpu_reference_step(O, x, rows, q, k, threads):
    """What the reference does per filtered query (ref vector_database.py:508-514):
    gather admissible rows into a temporary IndexFlatIP, then search it.  faiss
    parallelises over queries only, so `threads` concurrent queries use all cores."""
    import concurrent.futures as cf

    def one(i):
        return O.search_gathered(x, rows, q[i:i + 1], k, nthreads=1)

    with cf.ThreadPoolExecutor(max_workers=threads) as ex:
        list(ex.map(one, range(q.shape[0])))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    O, x, rows = cpu_reference_setup(N_ROWS, DIM)
    q = O.synth_rows(SEED_Q, 0, threads * (args.steps + args.warmup), DIM)
    O.normalize_L2(q)
    for w in range(args.warmup):
        cpu_reference_step(O, x, rows, q[w * threads:(w + 1) * threads], TOPK, threads)
    t0 = time.perf_counter()
    for s in range(args.steps):
        a = (args.warmup + s) * threads
        cpu_reference_step(O, x, rows, q[a:a + threads], TOPK, threads)
    dt = time.perf_counter() - t0
    qps = args.steps * threads / dt
    sample = f"{args.steps} steps x {threads} concurrent single-thread queries (gather {len(rows)} rows + scan), full 1M x 384"
    line = {
        "impl": "reference", "metric": METRIC, "value": qps, "unit": "queries/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "rows_per_gpu": N_ROWS, "dim": DIM, "k": TOPK,
                   "note": "CPU arm: oracle port of faiss IndexFlatIP (faiss-cpu not installable here); "
                           "host cores only, 1M rows regardless of --gpus"},
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    import minivectordb_b200 as mv
    from minivectordb_b200 import _native, synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    K, W = args.steps, args.warmup
    n, d, k = args.rows, args.dim, TOPK
    custom = (n, d) != (N_ROWS, DIM) or args.no_filter
    workload = WORKLOAD if not custom else (
        f"custom: {n} x {d} fp32 per GPU, nq=1, k={k}, " + ("no filter" if args.no_filter else "~50% filter bitmask"))
    peak, peak_src = measured_peak()

    from minivectordb_b200.distributed import RowShardedIndex
    index = RowShardedIndex(d, device=local, exchange=args.exchange)   # one row shard per rank, resident in HBM
    index.add(synthetic=(SEED_DB, rank * n, n, 0), normalize=True)
    eng = index.engine
    ld = eng.device_view()[1]
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

    def step(i, q_t=None, m_t=None):
        # scan this rank's shard (+ for N>1: NCCL all-gather of k (score,label) pairs and merge on every rank)
        if args.no_filter:
            return index.search_device(q_t if q_t is not None else q_dev[i:i + 1], k)
        return index.search_device(q_t if q_t is not None else q_dev[i:i + 1], k,
                                   m_t if m_t is not None else mask_dev, n)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

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
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    launches0 = _native.launch_count()
    e_begin, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e_begin.record(stream)
    for i in range(K):
        step(W + i)
    e_end.record(stream)
    barrier()
    launches = _native.launch_count() - launches0
    total_s = e_begin.elapsed_time(e_end) * 1e-3
    t = torch.tensor([total_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_s = float(t.item())
    # per-launch durations (isolated launches: the event between two searches serialises them)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
    barrier()
    ev[0].record(stream)
    for i in range(K):
        step(W + i)
        ev[i + 1].record(stream)
    barrier()
    per_step = np.array([ev[i].elapsed_time(ev[i + 1]) * 1e-3 for i in range(K)])
    eng.set_option("pdl", 0)

    # ---- end-to-end leg: host buffers through the C ABI (mvdb_index_search) ----
    # every step copies the query (d*4 B) and the packed filter (n/8 B) H2D and
    # the (D, I) result D2H inside the timed region.
    e2e_lat = []

    def e2e_step(i):
        if world == 1:
            if args.no_filter:
                return eng.search(q_host[i:i + 1], k)
            return eng.search(q_host[i:i + 1], k, mask=packed, mask_rows=n)
        # one pinned H2D ([filter words | query]), scan + fused exchange + merge, one D2H ([labels | distances])
        if args.no_filter:
            return index.search_packed(q_host[i:i + 1], k)
        return index.search_packed(q_host[i:i + 1], k, words, n)

    for i in range(W):
        e2e_step(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(K):
        a = time.perf_counter()
        e2e_step(W + i)
        e2e_lat.append(time.perf_counter() - a)
    barrier()
    e2e_total = time.perf_counter() - t0
    t = torch.tensor([e2e_total], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_total = float(t.item())
    clk = clocks.stop() if rank == 0 else None

    if rank == 0:
        alg_bytes = n * ld * 4 + (0 if args.no_filter else (n + 7) // 8)   # N*d*4 + ceil(N/8) with a filter mask
        kern_s = float(np.mean(per_step)) if world == 1 else None
        qps_global = K / total_s
        line = {
            "metric": METRIC, "value": qps_global * world, "unit": "queries/s", "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": total_s / K * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload, "rows_per_gpu": n, "rows_total": n * world, "dim": d, "k": k,
                       "filter_keep": None if args.no_filter else float(adm.mean()),
                       "l2": f"matrix {n * ld * 4 / 1e9:.2f} GB per GPU vs 126 MB L2; streamed with L2 evict_first, "
                             "distinct query each step (no flush needed above ~0.3 GB)",
                       "unit_note": "value = shard scans/s over all ranks = n_gpus x global QPS "
                                    f"(each rank scans its own {n} x {d} shard per query)",
                       "launch": "plain stream order" if args.no_pdl else
                                 "programmatic dependent launch: back-to-back searches overlap one search's merge tail with the next scan",
                       "parallelism": f"row-shard x{world}" + (f", exchange={index.exchange}" if world > 1 else "")},
            "qps_global": qps_global,
            "p50_latency_us": float(np.median(per_step) * 1e6),
            "e2e": {"value": K / e2e_total * world, "unit": "queries/s",
                    "h2d_bytes_per_step": d * 4 + (0 if args.no_filter else (n + 7) // 8),
                    "d2h_bytes_per_step": k * 12, "p50_latency_us": float(np.median(e2e_lat) * 1e6),
                    "qps_global": K / e2e_total,
                    "api": ("mvdb_index_search (C ABI, host buffers; H2D query+mask, D2H results inside)" if world == 1 else
                            "RowShardedIndex.search_packed: one pinned H2D (mask+query) -> scan + exchange + merge -> one D2H")},
            "gpu_launches": int(launches),
            "clocks": clk,
        }
        if world == 1:
            ach = alg_bytes / kern_s / 1e9
            ach_pipe = alg_bytes / (total_s / K) / 1e9
            line["roofline"] = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                                "traffic": None, "kernel": f"scan_q1_kernel<{(ld // 4 + 31) // 32},tma>",
                                "alg_bytes_per_launch": alg_bytes, "avg_launch_us": kern_s * 1e6,
                                "peak_source": peak_src,
                                "note": "achieved/frac use the ISOLATED per-launch duration (an event after every "
                                        "search); `value` is the back-to-back rate of the timed region, where "
                                        "consecutive launches overlap (pipelined_*). The denominator is a read+write "
                                        "copy peak, a read-only stream can exceed it.",
                                "pipelined_achieved": ach_pipe, "pipelined_frac": ach_pipe / peak}
            try:
                if not custom:
                    prof = json.load(open(os.path.join(ROOT, "profiles", "scan_traffic.json")))
                    line["roofline"]["traffic"] = prof.get("dram_bytes_per_launch")
            except Exception:
                pass
            if not args.no_cpu_baseline and not custom:
                line["cpu_baseline"] = cpu_baseline_sample()
        print(json.dumps(line), flush=True)
    index.close()
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline_sample():
    """Oracle on the box's host cores, bounded sample of the same workload."""
    threads = os.cpu_count() or 1
    O, x, rows = cpu_reference_setup(N_ROWS, DIM)
    reps = 2
    q = O.synth_rows(SEED_Q, 0, threads * (reps + 1), DIM)
    O.normalize_L2(q)
    cpu_reference_step(O, x, rows, q[:threads], TOPK, threads)  # warm-up
    t0 = time.perf_counter()
    for r in range(reps):
        cpu_reference_step(O, x, rows, q[(r + 1) * threads:(r + 2) * threads], TOPK, threads)
    dt = time.perf_counter() - t0
    # the scan alone, single thread, as faiss runs one query (VDB:497)
    t1 = time.perf_counter()
    O.search_flat_ip(x, q[:1], TOPK, nthreads=1)
    scan_1t = time.perf_counter() - t1
    return {"value": reps * threads / dt, "unit": "queries/s", "cores": threads, "kind": "port",
            "sample": f"{reps * threads} filtered queries (gather ~{len(rows)} rows + scan each, "
                      f"{threads} concurrent single-thread queries) on the full 1M x 384 matrix",
            "unfiltered_scan_single_thread_s": scan_1t}


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
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU leg (profiling runs under ncu)")
    ap.add_argument("--rows", type=int, default=N_ROWS, help="rows per GPU (default: BASELINE config 2)")
    ap.add_argument("--dim", type=int, default=DIM)
    ap.add_argument("--no-filter", action="store_true", help="unfiltered queries (BASELINE config 4 shape)")
    ap.add_argument("--exchange", default="auto", choices=["auto", "fused", "nccl"])
    ap.add_argument("--no-pdl", action="store_true", help="plain stream-ordered launches in the timed region (A/B)")
    args = ap.parse_args()
    if args.impl == "reference":
        if args.steps > 20:
            args.steps = 20  # each step is `cores` full filtered queries (~0.5 s): keep the run in minutes
        args.warmup = min(args.warmup, 2)
        run_reference(args)
    else:
        args.warmup = max(args.warmup, 3)
        run_b200(args)


if __name__ == "__main__":
    main()
