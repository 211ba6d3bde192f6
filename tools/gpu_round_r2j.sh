#!/bin/bash
# round 2, call J: RMSD pass with 4 solver warps per tile warp, then call I (launch list + K1 capture)
mkdir -p gpurun_out
O=gpurun_out
timeout -k 5 900 python -m pytest tests/test_gpu_rmsd.py tests/test_gpu_cluster.py -q -x > $O/r2j_pytest.log 2>&1; echo "pytest exit $?" >> $O/r2j_pytest.log
tail -4 $O/r2j_pytest.log
if grep -q "pytest exit 124\|pytest exit 137" $O/r2j_pytest.log; then echo "HANG"; exit 1; fi
timeout -k 5 900 python tools/config5_rmsd.py --check-k 100 > $O/r2j_config5_1gpu.json 2> $O/r2j_config5_1gpu.err; tail -2 $O/r2j_config5_1gpu.err; cat $O/r2j_config5_1gpu.json
timeout -k 5 600 ncu --set full --clock-control none --import-source on -k regex:rmsd_tile_pass -s 3 -c 1 \
    -o $O/r2j_rmsd_tile_pass -f python tools/config5_rmsd.py --frames 2000000 --k 6 --templates 200 --check-k 0 > $O/r2j_ncu_rmsd.log 2>&1
tail -2 $O/r2j_ncu_rmsd.log
bash tools/gpu_round_r2i.sh
