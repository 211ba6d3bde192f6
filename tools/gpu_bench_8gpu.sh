#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r1c_bench_8gpu.json 2> gpurun_out/r1c_bench_8gpu.err
tail -c 700 gpurun_out/r1c_bench_8gpu.json; tail -3 gpurun_out/r1c_bench_8gpu.err
