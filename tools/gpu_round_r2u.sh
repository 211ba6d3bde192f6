#!/bin/bash
# round 2, call U: K3 batched converter (8 warps, 4 segments in flight) + look-ahead chain diagnostics
mkdir -p gpurun_out
O=gpurun_out
timeout -k 5 300 python -m pytest tests/test_gpu_libdistance.py tests/test_gpu_cluster.py -q -x > $O/r2u_pytest.log 2>&1; echo "pytest exit $?" >> $O/r2u_pytest.log
tail -4 $O/r2u_pytest.log
if ! grep -q "pytest exit 0" $O/r2u_pytest.log; then echo "PARITY FAILED / HANG, stopping"; grep -E "^E |Error|assert" $O/r2u_pytest.log | head -30; exit 1; fi
timeout -k 5 400 python tools/assign_points.py > $O/r2u_assign_points.log 2>&1; echo "exit $?" >> $O/r2u_assign_points.log
cat $O/r2u_assign_points.log
timeout -k 5 600 python tools/lookahead_diag.py > $O/r2u_lookahead_diag.log 2>&1; echo "exit $?" >> $O/r2u_lookahead_diag.log
cat $O/r2u_lookahead_diag.log
