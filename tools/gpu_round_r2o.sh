#!/bin/bash
# round 2, call O: final validation -- GPU suite, smoke, bench (ours + reference arm), launch list, ncu captures
mkdir -p gpurun_out
O=gpurun_out
timeout -k 5 1500 python -m pytest tests -m gpu -q > $O/r2o_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/r2o_pytest_gpu.log
tail -5 $O/r2o_pytest_gpu.log
timeout -k 5 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2o_smoke.log 2>&1; tail -1 $O/r2o_smoke.log
timeout 1500 python bench.py > $O/r2o_bench_1gpu.json 2> $O/r2o_bench_1gpu.err
tail -c 3500 $O/r2o_bench_1gpu.json; tail -3 $O/r2o_bench_1gpu.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/r2o_bench_reference.json 2>/dev/null
tail -c 400 $O/r2o_bench_reference.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:tica_|kcenters_|candidate_|rmsd_|assign_' -c 400 --csv --log-file $O/r2o_launches_step.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-f64-check --no-ref-schedule --no-other-configs > $O/r2o_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tica_umma_v2_kernel --launch-skip 1 --launch-count 1 \
   -o $O/r2o_k1_v2_full -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-f64-check --no-ref-schedule --no-other-configs > $O/r2o_ncu_k1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:kcenters_multi_pass --launch-skip 3 --launch-count 2 \
   -o $O/r2o_k2b_full -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-f64-check --no-ref-schedule --no-other-configs > $O/r2o_ncu_k2b.log 2>&1
timeout -k 5 600 ncu --set full --clock-control none --import-source on -k regex:rmsd_tile_pass -s 3 -c 1 \
    -o $O/r2o_rmsd_tile_pass -f python tools/config5_rmsd.py --frames 2000000 --k 6 --templates 200 --check-k 0 > $O/r2o_ncu_rmsd.log 2>&1
timeout -k 5 600 ncu --set full --clock-control none --import-source on -k regex:assign_umma_kernel -s 1 -c 1 \
    -o $O/r2o_assign_umma -f python tools/profile_assign.py > $O/r2o_ncu_assign.log 2>&1
timeout -k 5 300 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:assign_' -c 40 --csv --log-file $O/r2o_launches_assign.csv \
    python tools/profile_assign.py > $O/r2o_assign_times.log 2>&1
timeout -k 5 120 python tools/profile_assign.py > $O/r2o_assign_plain.log 2>&1; cat $O/r2o_assign_plain.log
ls -la $O/*.ncu-rep | tail -5
