#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r1l.log; : > $L
for mb in 2 3 4; do
MSMB200_K2B_MINB=$mb timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/r1l_bench_$mb.json 2> gpurun_out/r1l_bench_$mb.err
python - >> $L <<PY
import json
l=json.loads(open("gpurun_out/r1l_bench_$mb.json").read().strip().splitlines()[-1])
print("minb=$mb", l["value"], l["ms_per_step"], l["phases_ms"], l["roofline"]["ms_per_launch"], l["roofline"]["centres_per_launch"], l["clocks"])
PY
done
cat $L
