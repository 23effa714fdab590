# ncu evidence for the round (run under gpurun, ONE GPU): launch list of the default bench command, `--set full`
# of the fp32 scan on the C4 shard and on C2, and of the int8 shadow scan.  Numbers printed under ncu are never
# bench values.  Afterwards, here:  python tools/ncu_summarise.py   (writes profiles/scan_traffic.json + raw CSVs)
set -x
B="python bench.py --no-cpu-baseline --no-parity"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_bench_c4.csv $B --steps 10 --warmup 3 > gpurun_out/ncu_b1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:scan_q1 -s 5 -c 3 -o gpurun_out/r02_prof_scan_c4 $B --steps 6 --warmup 3 --no-secondary > gpurun_out/ncu_b2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:scan_q1 -s 5 -c 3 -o gpurun_out/r02_prof_scan_c2 $B --workload c2 --steps 6 --warmup 3 > gpurun_out/ncu_b3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:scan_i8 -s 5 -c 3 -o gpurun_out/r02_prof_scan_i8_c4 python tools/shadow_one.py > gpurun_out/ncu_b4.log 2>&1
# gpurun merges at most 64 MiB back: keep the raw pages as CSV, drop the reports
for r in r02_prof_scan_c4 r02_prof_scan_c2 r02_prof_scan_i8_c4; do
  ncu -i gpurun_out/$r.ncu-rep --page raw --csv > gpurun_out/$r.raw.csv 2>/dev/null
  ncu -i gpurun_out/$r.ncu-rep --page details --csv 2>/dev/null | head -400 > gpurun_out/$r.details.csv
  rm -f gpurun_out/$r.ncu-rep
done
ls -la gpurun_out/*.csv
