#!/bin/bash
# round 2, final validation: GPU suite, smoke, bench (ours + reference arm), launch list, ncu captures of the
# kernels this session changed (K1 MN-major mode, look-ahead passes, streamed K3)
mkdir -p gpurun_out
O=gpurun_out
timeout -k 5 1500 python -m pytest tests -m gpu -q > $O/r2f_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/r2f_pytest_gpu.log
tail -4 $O/r2f_pytest_gpu.log
timeout -k 5 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2f_smoke.log 2>&1; tail -1 $O/r2f_smoke.log
timeout 1500 python bench.py > $O/r2f_bench_1gpu.json 2> $O/r2f_bench_1gpu.err
tail -c 1500 $O/r2f_bench_1gpu.json; tail -3 $O/r2f_bench_1gpu.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/r2f_bench_reference.json 2>/dev/null
tail -c 400 $O/r2f_bench_reference.json
Q="--no-cpu-baseline --no-e2e --no-f64-check --no-ref-schedule --no-other-configs"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:tica_|kcenters_|candidate_|rmsd_|assign_' -c 400 --csv --log-file $O/r2f_launches_step.csv \
    python bench.py --steps 2 --warmup 1 $Q > $O/r2f_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tica_umma_v2_kernel --launch-skip 1 --launch-count 1 \
   -o $O/r2f_k1_mn_full -f python bench.py --steps 1 --warmup 1 $Q > $O/r2f_ncu_k1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:kcenters_first_pass|kcenters_fused_pass' --launch-skip 2 --launch-count 2 \
   -o $O/r2f_k2b_full -f python bench.py --steps 1 --warmup 1 $Q > $O/r2f_ncu_k2b.log 2>&1
timeout -k 5 600 ncu --set full --clock-control none --import-source on -k regex:assign_umma_stream_kernel -s 1 -c 1 \
    -o $O/r2f_assign_stream -f python tools/profile_assign.py 10000000 128 2000 > $O/r2f_ncu_assign.log 2>&1
timeout -k 5 300 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:assign_' -c 40 --csv --log-file $O/r2f_launches_assign_stream.csv \
    python tools/profile_assign.py 10000000 128 2000 > $O/r2f_assign_times.log 2>&1
ls -la $O/*.ncu-rep | tail -5
