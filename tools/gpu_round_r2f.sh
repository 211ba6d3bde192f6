#!/bin/bash
# round 2, call F: K1 with two N-wide regions + 8 drain warps; ncu of the RMSD tile pass
mkdir -p gpurun_out
O=gpurun_out
: > $O/r2f_v2_smoke.log
for shape in "256 3 4000" "128 6 5000" "64 6 5000" "160 5 3000" "96 5 3000 3" "32 4 3000" "256 2 700 37" "224 9 1500 1"; do
    timeout -k 5 90 python tools/v2_smoke.py $shape >> $O/r2f_v2_smoke.log 2>&1
    echo "exit $? for $shape" >> $O/r2f_v2_smoke.log
done
grep -v "^\[umma" $O/r2f_v2_smoke.log | tail -16
if grep -q "exit 124\|exit 137" $O/r2f_v2_smoke.log; then echo "HANG detected, stopping"; exit 1; fi
timeout -k 5 900 python -m pytest tests/test_gpu_tica.py tests/test_gpu_tica_at_size.py -q > $O/r2f_pytest_tica.log 2>&1; echo "pytest exit $?" >> $O/r2f_pytest_tica.log
tail -8 $O/r2f_pytest_tica.log
MSMB200_UMMA_DEBUG=1 timeout -k 5 600 python tools/k1_experiments.py --frames 8000000 \
    v2: v1:MSMB200_UMMA_V1=1 v2_nodrain:MSMB200_UMMA_DBGMODE=2 v2_noconv:MSMB200_UMMA_DBGMODE=1 > $O/r2f_k1_experiments.log 2>&1
grep -v "^\[umma" $O/r2f_k1_experiments.log | tail; grep "umma" $O/r2f_k1_experiments.log | awk 'NR%7==1' | tail -4
MSMB200_UMMA_DEBUG=1 timeout -k 5 600 python tools/k1_experiments.py --frames 10000000 --features 64 \
    v2: > $O/r2f_k1_experiments_d64.log 2>&1
grep -v "^\[umma" $O/r2f_k1_experiments_d64.log | tail -2; grep "umma v2 dbg" $O/r2f_k1_experiments_d64.log | tail -1
MSMB200_UMMA_DEBUG=1 timeout -k 5 600 python tools/k1_experiments.py --frames 10000000 --features 128 \
    v2: > $O/r2f_k1_experiments_d128.log 2>&1
grep -v "^\[umma" $O/r2f_k1_experiments_d128.log | tail -2; grep "umma v2 dbg" $O/r2f_k1_experiments_d128.log | tail -1
timeout -k 5 600 ncu --set full --clock-control none --import-source on -k regex:rmsd_tile_pass -s 3 -c 1 \
    -o $O/r2f_rmsd_tile_pass python tools/config5_rmsd.py --frames 2000000 --k 6 --templates 200 --check-k 0 > $O/r2f_ncu_rmsd.log 2>&1
tail -3 $O/r2f_ncu_rmsd.log
