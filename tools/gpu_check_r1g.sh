#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r1g.log; : > $L
timeout 900 python -m pytest tests/test_gpu_tica.py -m gpu -x -q 2>&1 | tail -3 >> $L
for col in 0 1; do
  echo "== collector=$col" >> $L
  MSMB200_UMMA_COLLECTOR=$col ENGINES=umma_3xf16,umma_6xbf16,umma_3xtf32 SLABS=16,32,64 NSEQ=40 \
    timeout 600 python tools/umma_accuracy.py 2>&1 | grep -v simt >> $L
done
echo "== debug" >> $L
MSMB200_UMMA_DEBUG=1 ENGINES=umma_3xf16 SLABS=32 NSEQ=40 timeout 600 python tools/umma_accuracy.py 2>&1 | grep dbg | tail -1 >> $L
MSMB200_UMMA_COLLECTOR=1 MSMB200_UMMA_DEBUG=1 ENGINES=umma_3xf16 SLABS=32 NSEQ=40 timeout 600 python tools/umma_accuracy.py 2>&1 | grep dbg | tail -1 >> $L
cat $L
