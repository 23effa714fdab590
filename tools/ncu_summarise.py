"""Turn the .ncu-rep files that tools/ncu_refresh.sh left in gpurun_out/ into the committed evidence:
profiles/r02_scan_*_full_raw.csv (ncu --page raw) and profiles/scan_traffic.json (DRAM bytes per launch, keyed
by bench workload, stamped with a hash of the kernel sources they were captured from)."""
import csv, hashlib, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
h = hashlib.sha256()
for f in ("scan.cuh", "select.cuh", "device_utils.cuh"):
    h.update(open(os.path.join(ROOT, "minivectordb_b200", "csrc", f), "rb").read())
sha = h.hexdigest()[:16]
out = {}
for key, rep, rows, dim, raw in (("c4", "r02_prof_scan_c4", 12_500_000, 512, "r02_scan_q1_c4_full_raw.csv"),
                                 ("c2", "r02_prof_scan_c2", 1_000_000, 384, "r02_scan_q1_c2_full_raw.csv"),
                                 ("c4_shadow", "r02_prof_scan_i8_c4", 12_500_000, 512, "r02_scan_i8_c4_full_raw.csv")):
    src = os.path.join(ROOT, "gpurun_out", rep + ".ncu-rep")
    csv_src = os.path.join(ROOT, "gpurun_out", rep + ".raw.csv")
    dst = os.path.join(ROOT, "profiles", raw)
    if os.path.exists(csv_src):          # tools/ncu_refresh.sh exports the raw page on the GPU box (the reports are too big to travel)
        import shutil
        shutil.copy(csv_src, dst)
    elif os.path.exists(src):
        with open(dst, "w") as f:
            subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], stdout=f, check=True)
    else:
        continue
    rws = list(csv.reader(open(dst)))
    hdr, units = rws[0], rws[1]
    idx = {k: i for i, k in enumerate(hdr)}
    def val(r, k):
        return float(r[idx[k]]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "ms": 1e3, "us": 1, "ns": 1e-3}.get(units[idx[k]], 1)
    body = rws[2:]
    rd = [val(r, "dram__bytes_read.sum") for r in body]
    wr = [val(r, "dram__bytes_write.sum") for r in body]
    out[key] = {"rows": rows, "dim": dim, "kernel": body[0][idx["Kernel Name"]],
                "dram_bytes_per_launch": int(sum(rd) / len(rd) + sum(wr) / len(wr)), "dram_bytes_read": int(sum(rd) / len(rd)),
                "dram_bytes_write": int(sum(wr) / len(wr)), "launches_profiled": len(body),
                "gpu__dram_throughput_pct": [float(r[idx["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]]) for r in body],
                "gpu__time_duration_us": [round(val(r, "gpu__time_duration.sum"), 2) for r in body],
                "registers_per_thread": int(float(body[0][idx["launch__registers_per_thread"]])),
                "source": f"profiles/{raw} (ncu --set full --clock-control none, 3 launches; tools/ncu_refresh.sh)",
                "kernel_sources_sha256": sha}
json.dump(out, open(os.path.join(ROOT, "profiles", "scan_traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
