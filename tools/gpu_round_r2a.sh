#!/bin/bash
# round 2, call A: state of the tree after the parity / housekeeping work + drain experiments for K1
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > $O/r2a_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/r2a_pytest_gpu.log
tail -5 $O/r2a_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2a_smoke.log 2>&1; tail -1 $O/r2a_smoke.log
MSMB200_UMMA_DEBUG=1 timeout 600 python tools/k1_experiments.py --frames 8000000 \
    red2:MSMB200_UMMA_FLUSH_RED=2 red3_bulk:MSMB200_UMMA_FLUSH_RED=3 red4_v4:MSMB200_UMMA_FLUSH_RED=4 \
    red1_f64:MSMB200_UMMA_FLUSH_RED=1 nodrain:MSMB200_UMMA_DBGMODE=2 noconv:MSMB200_UMMA_DBGMODE=1 \
    onemma:MSMB200_UMMA_DBGMODE=4 > $O/r2a_k1_experiments.log 2>&1
cat $O/r2a_k1_experiments.log | tail -30
timeout 1500 python bench.py > $O/r2a_bench_1gpu.json 2> $O/r2a_bench_1gpu.err
tail -c 3000 $O/r2a_bench_1gpu.json
tail -5 $O/r2a_bench_1gpu.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/r2a_bench_reference.json 2>/dev/null
tail -c 600 $O/r2a_bench_reference.json
