#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r1i.log; : > $L
timeout 900 python -m pytest tests/test_gpu_tica.py -m gpu -x -q 2>&1 | tail -3 >> $L
for fix in 0 150 210 260; do
  echo "== biasfix=$fix" >> $L
  MSMB200_UMMA_BIASFIX=$fix ENGINES=umma_3xf16,umma_6xbf16 SLABS=32,64,128,256 NSEQ=40 \
    timeout 600 python tools/umma_accuracy.py 2>&1 | grep -v simt >> $L
done
echo "== debug" >> $L
MSMB200_UMMA_DEBUG=1 ENGINES=umma_3xf16 SLABS=32 NSEQ=40 timeout 600 python tools/umma_accuracy.py 2>&1 | grep dbg | tail -1 >> $L
cat $L
