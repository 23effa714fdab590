set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for N in 8 4 2; do
  timeout 300 $TR --nproc-per-node $N --master-port 2951$N bench.py --gpus $N --steps 200 --warmup 10 > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
  cat gpurun_out/scale_n$N.json; tail -2 gpurun_out/scale_n$N.err
done
timeout 200 python bench.py --no-cpu-baseline > gpurun_out/scale_n1.json 2>/dev/null; cat gpurun_out/scale_n1.json
# NCCL transport at 8 for comparison
timeout 300 $TR --nproc-per-node 8 --master-port 29531 bench.py --gpus 8 --steps 200 --warmup 10 --exchange nccl > gpurun_out/scale_n8_nccl.json 2>/dev/null; cat gpurun_out/scale_n8_nccl.json
# BASELINE config 4: 100M x 512 over 8 GPUs (12.5M rows per GPU), unfiltered single query
timeout 400 $TR --nproc-per-node 8 --master-port 29541 bench.py --gpus 8 --steps 100 --warmup 5 --rows 12500000 --dim 512 --no-filter > gpurun_out/c4_n8.json 2> gpurun_out/c4_n8.err; cat gpurun_out/c4_n8.json; tail -2 gpurun_out/c4_n8.err
