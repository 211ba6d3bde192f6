#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r1j.log; : > $L
for hb in 11 10 9 8; do
  echo "== hbits=$hb" >> $L
  MSMB200_UMMA_HBITS=$hb ENGINES=umma_3xf16 SLABS=32,64,128 NSEQ=40 \
    timeout 600 python tools/umma_accuracy.py 2>&1 | grep -v simt >> $L
done
cat $L
