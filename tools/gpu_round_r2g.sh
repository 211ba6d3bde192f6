#!/bin/bash
# round 2, call G: whole GPU suite, config 5 with the pipelined RMSD pass, narrow-D K1, bench
mkdir -p gpurun_out
O=gpurun_out
timeout -k 5 1500 python -m pytest tests -m gpu -q > $O/r2g_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/r2g_pytest_gpu.log
tail -12 $O/r2g_pytest_gpu.log
timeout -k 5 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2g_smoke.log 2>&1; tail -1 $O/r2g_smoke.log
MSMB200_UMMA_DEBUG=1 timeout -k 5 600 python tools/k1_experiments.py --frames 10000000 --features 64 \
    v2: > $O/r2g_k1_experiments_d64.log 2>&1
grep -v "^\[umma" $O/r2g_k1_experiments_d64.log | tail -2; grep "umma v2 dbg" $O/r2g_k1_experiments_d64.log | tail -1
timeout -k 5 600 python tools/config5_rmsd.py --frames 500000 --k 200 --templates 200 --check-k 200 > $O/r2g_config5_small.json 2> $O/r2g_config5_small.err; tail -2 $O/r2g_config5_small.err; cat $O/r2g_config5_small.json
timeout -k 5 900 python tools/config5_rmsd.py --check-k 100 > $O/r2g_config5_1gpu.json 2> $O/r2g_config5_1gpu.err; tail -2 $O/r2g_config5_1gpu.err; cat $O/r2g_config5_1gpu.json
timeout 1500 python bench.py > $O/r2g_bench_1gpu.json 2> $O/r2g_bench_1gpu.err
tail -c 2500 $O/r2g_bench_1gpu.json
tail -5 $O/r2g_bench_1gpu.err
