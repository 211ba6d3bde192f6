#!/bin/bash
# round 2, call AA: K1 MN-major rolling-window mode (MSMB200_UMMA_MN=1): first contact, parity, timing
mkdir -p gpurun_out
O=gpurun_out
for shape in "256 3 4000 10" "128 6 5000 10" "256 2 700 32" "256 5 333 1" "128 3 100 7"; do
    MSMB200_UMMA_MN=1 timeout -k 5 90 python tools/v2_smoke.py $shape >> $O/r2aa_smoke.log 2>&1
    echo "exit $? for $shape" >> $O/r2aa_smoke.log
done
grep -v "^\[umma" $O/r2aa_smoke.log | tail -12
if grep -q "exit 124\|exit 137" $O/r2aa_smoke.log; then echo "HANG detected, stopping"; exit 1; fi
if grep -q "MISMATCH\|Error\|error" $O/r2aa_smoke.log; then echo "MISMATCH, stopping after the timing run"; fi
MSMB200_UMMA_DEBUG=1 timeout -k 5 300 python tools/k1_experiments.py --frames 8000000 v2: mn:MSMB200_UMMA_MN=1 > $O/r2aa_k1.log 2>&1
grep -v "^\[umma" $O/r2aa_k1.log | tail -3; grep "umma" $O/r2aa_k1.log | awk 'NR%7==1' | tail -2
if grep -q "MISMATCH" $O/r2aa_smoke.log; then exit 1; fi
MSMB200_UMMA_MN=1 timeout -k 5 600 python -m pytest tests/test_gpu_tica.py tests/test_gpu_tica_at_size.py -q -x > $O/r2aa_pytest.log 2>&1; echo "pytest exit $?" >> $O/r2aa_pytest.log
tail -3 $O/r2aa_pytest.log
MSMB200_UMMA_DEBUG=1 timeout -k 5 300 python tools/k1_experiments.py --frames 10000000 --features 128 v2: mn:MSMB200_UMMA_MN=1 > $O/r2aa_k1_d128.log 2>&1
grep -v "^\[umma" $O/r2aa_k1_d128.log | tail -2; grep "umma" $O/r2aa_k1_d128.log | awk 'NR%7==1' | tail -2
