#!/bin/bash
# final validation of the round-1 kernels: GPU tests, bench lines, ncu launch list + full captures
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q > $O/r1c_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/r1c_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/r1c_smoke.log 2>&1; tail -1 $O/r1c_smoke.log
tail -3 $O/r1c_pytest_gpu.log
timeout 1200 python bench.py > $O/r1c_bench_1gpu.json 2> $O/r1c_bench_1gpu.err
tail -c 1500 $O/r1c_bench_1gpu.json
timeout 900 python bench.py --no-lookahead --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $O/r1c_bench_1gpu_no_lookahead.json 2>/dev/null
timeout 900 python bench.py --engine umma_6xbf16 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $O/r1c_bench_1gpu_6xbf16.json 2>/dev/null
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $O/r1c_bench_reference.json 2>/dev/null
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:tica_|kcenters_|candidate_' -c 400 --csv --log-file $O/r1c_launches_step.csv \
   python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > $O/r1c_launches_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:kcenters_multi_pass --launch-skip 3 --launch-count 2 \
   -o $O/r1c_k2b_full -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $O/r1c_ncu_k2b.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tica_umma_kernel --launch-skip 2 --launch-count 1 \
   -o $O/r1c_k1_full -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $O/r1c_ncu_k1.log 2>&1
ls -la $O/*.ncu-rep
