#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r1e.log; : > $L
for mode in 0 1 2 3 4 5 7; do
  echo "== dbgmode=$mode" >> $L
  MSMB200_UMMA_DBGMODE=$mode MSMB200_UMMA_DEBUG=1 ENGINES=umma_3xf16 SLABS=32 NSEQ=40 timeout 600 python tools/umma_accuracy.py 2>&1 | grep "dbg\|umma_3xf16" | tail -2 >> $L
done
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw --format=csv >> $L
cat $L
