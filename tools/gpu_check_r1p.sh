#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:kcenters_multi_pass --launch-skip 4 --launch-count 1 \
   -o gpurun_out/r1p_k2b_fused -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r1p_ncu.log 2>&1
tail -2 gpurun_out/r1p_ncu.log
