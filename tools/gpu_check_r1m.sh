#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r1m.log; : > $L
timeout 900 python -m pytest tests/test_gpu_lookahead.py -m gpu -x -q 2>&1 | tail -3 >> $L
timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/r1m_bench.json 2> gpurun_out/r1m_bench.err
python - >> $L <<PY
import json
l=json.loads(open("gpurun_out/r1m_bench.json").read().strip().splitlines()[-1])
print(l["value"], l["ms_per_step"], l["phases_ms"], l["roofline"]["ms_per_launch"], l["clocks"])
PY
tail -3 gpurun_out/r1m_bench.err >> $L
cat $L
