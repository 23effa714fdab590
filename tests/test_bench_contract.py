"""CPU: the reference arm of bench.py (`--impl reference`) runs without a GPU; check that it prints
exactly one JSON line on stdout with the keys the driver's contract names."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    # MVDB_BENCH_CPU_ROWS keeps the host-resident sample small here (the default workload is 12.5M x 512 = 25.6 GB)
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT, env={**os.environ, "MVDB_BENCH_CPU_ROWS": "200000"})
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["unit"] == "queries/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["steps"] == 1 and d["warmup"] == 0 and d["vs_baseline"] is None
    assert "workload" in d["config"] and d["config"]["rows_per_gpu"] == 12_500_000 and d["config"]["dim"] == 512
    # the two arms of one (workload, N) must print the SAME config object (the driver compares them)
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"] == bench.make_config("c4", 1)
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["sample"] and cb["value"] == d["value"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


def test_parity_key_logic_on_cpu():
    """bench.parity_check (streamed oracle + cross-leg / cross-rank comparison) accepts the right answer
    and flags a wrong one -- here the "GPU" answers are produced by the oracle on the whole matrix."""
    import numpy as np
    sys.path.insert(0, ROOT)
    import bench
    from oracle import oracle as O
    n, d, k = 50_000, 64, 10
    old = bench.ORACLE_CHUNK
    bench.ORACLE_CHUNK = 8192   # several chunks
    try:
        x = O.synth_rows(bench.SEED_DB, 0, n, d)
        O.normalize_L2(x)
        q = O.synth_rows(bench.SEED_Q, 0, 4, d)
        O.normalize_L2(q)
        res = [O.search_flat_ip(x, q[i:i + 1], k) for i in range(4)]
        rep = bench.parity_check(0, 1, n, d, k, None, q, res, res)
        assert rep["ok"] and rep["ids_equal"] == 40 and rep["rows_checked"] == n and rep["real_errors"] == 0
        bad = [(D.copy(), I.copy()) for D, I in res]
        bad[2][1][0, 3] = 7   # a wrong id
        rep = bench.parity_check(0, 1, n, d, k, None, q, bad, res)
        assert not rep["ok"] and rep["real_errors"] >= 1 and not rep["device_and_e2e_legs_identical"]
        adm = np.random.default_rng(0).random(n) < 0.5
        resm = [O.search_masked(x, adm, q[i:i + 1], k) for i in range(4)]
        assert bench.parity_check(0, 1, n, d, k, adm, q, resm, resm)["ok"]
    finally:
        bench.ORACLE_CHUNK = old
