#!/bin/bash
mkdir -p gpurun_out
ENGINES=umma_3xf16,umma_6xbf16,umma_3xbf16 SLABS=8,16,32,64,128 NSEQ=40 timeout 600 python tools/umma_accuracy.py > gpurun_out/r1c_accuracy.log 2>&1
cat gpurun_out/r1c_accuracy.log
