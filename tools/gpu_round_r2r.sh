#!/bin/bash
# round 2, call R: direct-load converter experiment (MSMB200_UMMA_DIRECT=1)
mkdir -p gpurun_out
O=gpurun_out
for shape in "256 3 4000" "128 6 5000" "256 2 700 37" "224 9 1500 1"; do
    MSMB200_UMMA_DIRECT=1 timeout -k 5 90 python tools/v2_smoke.py $shape >> $O/r2r_smoke.log 2>&1
    echo "exit $? for $shape" >> $O/r2r_smoke.log
done
grep -v "^\[umma" $O/r2r_smoke.log | tail -8
if grep -q "exit 124\|exit 137" $O/r2r_smoke.log; then echo "HANG detected, stopping"; exit 1; fi
MSMB200_UMMA_DIRECT=1 timeout -k 5 600 python -m pytest tests/test_gpu_tica.py -q -x > $O/r2r_pytest.log 2>&1; echo "pytest exit $?" >> $O/r2r_pytest.log
tail -3 $O/r2r_pytest.log
MSMB200_UMMA_DEBUG=1 timeout -k 5 600 python tools/k1_experiments.py --frames 8000000 v2: direct:MSMB200_UMMA_DIRECT=1 > $O/r2r_k1.log 2>&1
grep -v "^\[umma" $O/r2r_k1.log | tail -3; grep "umma" $O/r2r_k1.log | awk 'NR%7==1' | tail -2
MSMB200_UMMA_DEBUG=1 timeout -k 5 600 python tools/k1_experiments.py --frames 10000000 --features 128 v2: direct:MSMB200_UMMA_DIRECT=1 > $O/r2r_k1_d128.log 2>&1
grep -v "^\[umma" $O/r2r_k1_d128.log | tail -2; grep "umma" $O/r2r_k1_d128.log | awk 'NR%7==1' | tail -2
timeout -k 5 300 python -m pytest tests/test_gpu_widen.py -q -x -k "regular_spatial" > $O/r2r_pytest_rs.log 2>&1; tail -3 $O/r2r_pytest_rs.log
timeout -k 5 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2r_smoke2.log 2>&1; tail -2 $O/r2r_smoke2.log
