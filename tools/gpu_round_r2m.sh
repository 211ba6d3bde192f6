#!/bin/bash
# round 2, call M: branch-free converter fast path, grid-stride float64 rescue
mkdir -p gpurun_out
O=gpurun_out
timeout -k 5 900 python -m pytest tests/test_gpu_tica.py tests/test_gpu_tica_at_size.py tests/test_gpu_widen.py -q -x > $O/r2m1_pytest.log 2>&1; echo "pytest exit $?" >> $O/r2m1_pytest.log
tail -4 $O/r2m1_pytest.log
if grep -q "pytest exit 124\|pytest exit 137" $O/r2m1_pytest.log; then echo "HANG"; exit 1; fi
MSMB200_UMMA_DEBUG=1 timeout -k 5 600 python tools/k1_experiments.py --frames 8000000 v2: v1:MSMB200_UMMA_V1=1 f64:MSMB200_UMMA_V1=0 > $O/r2m1_k1_experiments.log 2>&1
grep -v "^\[umma" $O/r2m1_k1_experiments.log | tail -4; grep "umma" $O/r2m1_k1_experiments.log | awk 'NR%7==1' | head -1
MSMB200_UMMA_DEBUG=1 timeout -k 5 600 python tools/k1_experiments.py --frames 10000000 --features 128 v2: > $O/r2m1_k1_experiments_d128.log 2>&1
grep -v "^\[umma" $O/r2m1_k1_experiments_d128.log | tail -1; grep "umma v2 dbg" $O/r2m1_k1_experiments_d128.log | tail -1
timeout -k 5 300 python tools/k1_experiments.py --frames 4000000 --engine simt_f64 --reps 2 f64: > $O/r2m1_f64.log 2>&1; tail -1 $O/r2m1_f64.log
