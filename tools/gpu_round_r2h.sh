#!/bin/bash
# round 2, call H: RMSD pass with solver warps
mkdir -p gpurun_out
O=gpurun_out
timeout -k 5 900 python -m pytest tests/test_gpu_rmsd.py tests/test_gpu_cluster.py tests/test_gpu_lookahead.py tests/test_gpu_agglomerative.py -q -x > $O/r2h_pytest.log 2>&1; echo "pytest exit $?" >> $O/r2h_pytest.log
tail -6 $O/r2h_pytest.log
if grep -q "pytest exit 124\|pytest exit 137" $O/r2h_pytest.log; then echo "HANG"; exit 1; fi
timeout -k 5 600 python tools/config5_rmsd.py --frames 500000 --k 200 --templates 200 --check-k 200 > $O/r2h_config5_small.json 2> $O/r2h_config5_small.err; tail -2 $O/r2h_config5_small.err; cat $O/r2h_config5_small.json
timeout -k 5 900 python tools/config5_rmsd.py --check-k 100 > $O/r2h_config5_1gpu.json 2> $O/r2h_config5_1gpu.err; tail -2 $O/r2h_config5_1gpu.err; cat $O/r2h_config5_1gpu.json
timeout -k 5 600 ncu --set full --clock-control none --import-source on -k regex:rmsd_tile_pass -s 3 -c 1 \
    -o $O/r2h_rmsd_tile_pass python tools/config5_rmsd.py --frames 2000000 --k 6 --templates 200 --check-k 0 > $O/r2h_ncu_rmsd.log 2>&1
tail -2 $O/r2h_ncu_rmsd.log
