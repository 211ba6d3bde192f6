#!/bin/bash
mkdir -p gpurun_out
ENGINES=umma_3xf16 SLABS=32 NSEQ=40 timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:tica_umma_kernel --launch-count 1 -o gpurun_out/r1h_umma_f16 -f python tools/umma_accuracy.py > gpurun_out/r1h_ncu.log 2>&1
tail -2 gpurun_out/r1h_ncu.log
