#!/bin/bash
# round 2, call AK: bench.py --workload assign (config 3) on one GPU
mkdir -p gpurun_out
timeout 600 python bench.py --workload assign > gpurun_out/r2p_bench_assign_1gpu.json 2> gpurun_out/r2p_bench_assign_1gpu.err
tail -c 2500 gpurun_out/r2p_bench_assign_1gpu.json; tail -3 gpurun_out/r2p_bench_assign_1gpu.err
