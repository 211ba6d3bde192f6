#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r1d.log; : > $L
timeout 900 python -m pytest tests/test_gpu_tica.py -m gpu -x -q 2>&1 | tail -3 >> $L
for red in 1 0; do for col in 0 1; do
  echo "== flush_red=$red collector=$col" >> $L
  MSMB200_UMMA_FLUSH_RED=$red MSMB200_UMMA_COLLECTOR=$col ENGINES=umma_3xf16,umma_6xbf16 SLABS=16,32,64 NSEQ=40 \
    timeout 600 python tools/umma_accuracy.py 2>&1 | grep -v simt >> $L
done; done
echo "== debug red=1" >> $L
MSMB200_UMMA_DEBUG=1 ENGINES=umma_3xf16 SLABS=32 NSEQ=40 timeout 600 python tools/umma_accuracy.py 2>&1 | grep dbg | tail -1 >> $L
echo "== debug red=0" >> $L
MSMB200_UMMA_FLUSH_RED=0 MSMB200_UMMA_DEBUG=1 ENGINES=umma_3xf16 SLABS=32 NSEQ=40 timeout 600 python tools/umma_accuracy.py 2>&1 | grep dbg | tail -1 >> $L
cat $L
