#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python bench.py > $O/r1c_bench_1gpu.json 2> $O/r1c_bench_1gpu.err
tail -c 600 $O/r1c_bench_1gpu.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:tica_|kcenters_|candidate_' -c 400 --csv --log-file $O/r1c_launches_step.csv \
   python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > $O/r1c_launches_bench.log 2>&1
