#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r1k.log; : > $L
timeout 900 python -m pytest tests/test_gpu_lookahead.py tests/test_gpu_cluster.py -m gpu -x -q 2>&1 | tail -15 >> $L
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r1k_bench.json 2> gpurun_out/r1k_bench.err
tail -c 2500 gpurun_out/r1k_bench.json >> $L; tail -5 gpurun_out/r1k_bench.err >> $L
cat $L
