#!/bin/bash
# round 2, call T: K3 with streamed centre chunks (assign_umma_stream_kernel): parity, then timing points
mkdir -p gpurun_out
O=gpurun_out
timeout -k 5 300 python -m pytest tests/test_gpu_libdistance.py -q -x -k "streamed or tensor_core" > $O/r2t_pytest.log 2>&1; echo "pytest exit $?" >> $O/r2t_pytest.log
tail -4 $O/r2t_pytest.log
if ! grep -q "pytest exit 0" $O/r2t_pytest.log; then echo "PARITY FAILED / HANG, stopping"; grep -E "^E |Error|assert" $O/r2t_pytest.log | head -30; exit 1; fi
timeout -k 5 400 python tools/assign_points.py > $O/r2t_assign_points.log 2>&1; echo "exit $?" >> $O/r2t_assign_points.log
cat $O/r2t_assign_points.log
timeout -k 5 300 python -m pytest tests/test_gpu_libdistance.py tests/test_gpu_cluster.py tests/test_gpu_agglomerative.py -q -x > $O/r2t_pytest_all.log 2>&1; echo "pytest exit $?" >> $O/r2t_pytest_all.log
tail -3 $O/r2t_pytest_all.log
