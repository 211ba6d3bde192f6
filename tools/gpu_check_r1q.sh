#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r1q.log; : > $L
timeout 900 python -m pytest tests/test_gpu_tica.py -m gpu -x -q 2>&1 | tail -3 >> $L
for red in 2 1; do
  echo "== flush_red=$red" >> $L
  MSMB200_UMMA_FLUSH_RED=$red ENGINES=umma_3xf16,umma_6xbf16 SLABS=16,32,64 NSEQ=40 \
    timeout 600 python tools/umma_accuracy.py 2>&1 | grep -v simt >> $L
done
for fe in 4 16 64; do
echo "== debug red=2 fold_every=$fe" >> $L
MSMB200_UMMA_FOLD_EVERY=$fe MSMB200_UMMA_DEBUG=1 ENGINES=umma_3xf16 SLABS=32 NSEQ=40 timeout 600 python tools/umma_accuracy.py 2>&1 | grep "dbg\|umma_3xf16" | tail -2 >> $L
done
timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/r1q_bench.json 2> gpurun_out/r1q_bench.err
python - >> $L <<PY
import json
l=json.loads(open("gpurun_out/r1q_bench.json").read().strip().splitlines()[-1])
print(l["value"], l["ms_per_step"], l["phases_ms"], l["clocks"], l["check"])
PY
cat $L
