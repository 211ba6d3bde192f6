#!/bin/bash
# round 2, call N: K3 on the tensor cores
mkdir -p gpurun_out
O=gpurun_out
timeout -k 5 300 python -m pytest tests/test_gpu_libdistance.py -q -x -k "tensor_core or assign" > $O/r2n_pytest.log 2>&1; echo "pytest exit $?" >> $O/r2n_pytest.log
tail -25 $O/r2n_pytest.log
if grep -q "pytest exit 124\|pytest exit 137" $O/r2n_pytest.log; then echo "HANG"; exit 1; fi
timeout -k 5 600 python tools/config_checks.py > $O/r2n_config_checks.log 2>&1; tail -12 $O/r2n_config_checks.log
