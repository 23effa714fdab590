# ncu evidence for the round (one GPU): launch list of the bench command, full set of the scan
# kernel and of the default GEMM kernel.  Numbers printed by runs under ncu are never bench values.
set -x
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_q1 -s 5 -c 3 -f -o gpurun_out/prof_scan \
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ncu -i gpurun_out/prof_scan.ncu-rep --page raw --csv > gpurun_out/prof_scan_raw.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_topk_kernel_mc -s 7 -c 1 -f -o gpurun_out/prof_gemm \
    python tools/batched_one.py 2000000 1024 4096 100 2 1 1 > gpurun_out/ncu_gemm.log 2>&1
ncu -i gpurun_out/prof_gemm.ncu-rep --page raw --csv > gpurun_out/prof_gemm_raw.csv
ls -la gpurun_out/*.csv
