#!/bin/bash
mkdir -p gpurun_out
for mode in 0 1; do
MSMB200_UMMA_DBGMODE=$mode ENGINES=umma_3xf16 SLABS=32 NSEQ=40 timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:tica_umma_kernel --launch-count 1 -o gpurun_out/r1f_umma_f16_mode$mode -f python tools/umma_accuracy.py > gpurun_out/r1f_ncu_mode$mode.log 2>&1
tail -3 gpurun_out/r1f_ncu_mode$mode.log
done
ls -la gpurun_out/*.ncu-rep
