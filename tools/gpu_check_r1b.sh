#!/bin/bash
# one gpurun call: GPU test suite, accuracy/speed of the fp16 tICA engine, short bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/r1b_smi.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r1b_pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/r1b_pytest_gpu.log
tail -3 gpurun_out/r1b_pytest_gpu.log
for col in 0 1; do
  echo "== collector=$col" >> gpurun_out/r1b_accuracy.log
  MSMB200_UMMA_COLLECTOR=$col ENGINES=umma_3xf16,umma_3xbf16,umma_6xbf16 SLABS=64 NSEQ=40 \
    timeout 600 python tools/umma_accuracy.py >> gpurun_out/r1b_accuracy.log 2>&1
done
MSMB200_UMMA_DEBUG=1 ENGINES=umma_3xf16 SLABS=64 NSEQ=40 timeout 600 python tools/umma_accuracy.py >> gpurun_out/r1b_accuracy.log 2>&1
cat gpurun_out/r1b_accuracy.log
for col in 0 1; do
  MSMB200_UMMA_COLLECTOR=$col timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r1b_bench_col$col.json 2> gpurun_out/r1b_bench_col$col.err
  python - <<PY
import json
l=json.loads(open("gpurun_out/r1b_bench_col$col.json").read().strip().splitlines()[-1])
print("collector=$col", l["value"], l["ms_per_step"], l["phases_ms"], l["clocks"], l["e2e"]["value"])
PY
done
