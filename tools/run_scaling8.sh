# N = 8 only: weak scaling of config 2 (fused exchange, pipelined launches) and BASELINE config 4
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 200 $TR --nproc-per-node 8 --master-port 29518 bench.py --gpus 8 --steps 200 --warmup 10 > gpurun_out/scale_n8.json 2> gpurun_out/scale_n8.err
cat gpurun_out/scale_n8.json; tail -2 gpurun_out/scale_n8.err
timeout 300 $TR --nproc-per-node 8 --master-port 29541 bench.py --gpus 8 --steps 100 --warmup 5 --rows 12500000 --dim 512 --no-filter > gpurun_out/c4_n8.json 2> gpurun_out/c4_n8.err
cat gpurun_out/c4_n8.json; tail -2 gpurun_out/c4_n8.err
