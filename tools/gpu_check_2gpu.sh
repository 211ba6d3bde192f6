#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parallel.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r1n_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r1n_bench_2gpu.json 2> gpurun_out/r1n_bench_2gpu.err
tail -c 2500 gpurun_out/r1n_bench_2gpu.json >> gpurun_out/r1n_2gpu.log; tail -5 gpurun_out/r1n_bench_2gpu.err >> gpurun_out/r1n_2gpu.log
cat gpurun_out/r1n_2gpu.log
