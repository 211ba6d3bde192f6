#!/bin/bash
# full validation of the round-1b kernels: GPU tests, bench line, ncu launch list + full captures
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r1b_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r1b_pytest_gpu.log
tail -3 gpurun_out/r1b_pytest_gpu.log
timeout 1200 python bench.py > gpurun_out/r1b_bench_1gpu.json 2> gpurun_out/r1b_bench_1gpu.err
tail -c 3000 gpurun_out/r1b_bench_1gpu.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1b_launches_step.csv \
   python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r1b_launches_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:kcenters_pass_fast --launch-skip 3 --launch-count 1 \
   -o gpurun_out/r1b_k2_full -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r1b_ncu_k2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tica_umma_kernel --launch-skip 1 --launch-count 1 \
   -o gpurun_out/r1b_k1_full -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r1b_ncu_k1.log 2>&1
ls -la gpurun_out/*.ncu-rep
