#!/bin/bash
# round 2, call AC: stage count as a template parameter (K-major narrow kernel back to 1.9 ms?), full bench
mkdir -p gpurun_out
O=gpurun_out
timeout -k 5 900 python -m pytest tests -m gpu -q -x > $O/r2ac_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/r2ac_pytest_gpu.log
tail -3 $O/r2ac_pytest_gpu.log
MSMB200_UMMA_DEBUG=1 timeout -k 5 300 python tools/k1_experiments.py --frames 10000000 --features 64 v2: > $O/r2ac_k1_d64.log 2>&1
grep -v "^\[umma" $O/r2ac_k1_d64.log | tail -1; grep "umma v2 dbg" $O/r2ac_k1_d64.log | tail -1
MSMB200_UMMA_DEBUG=1 timeout -k 5 300 python tools/k1_experiments.py --frames 10000000 --features 128 kmajor:MSMB200_UMMA_MN=0 mn4:MSMB200_UMMA_MN_STAGES=4 mn5: > $O/r2ac_k1_d128.log 2>&1
grep -v "^\[umma" $O/r2ac_k1_d128.log | tail -3; grep "umma v2 dbg" $O/r2ac_k1_d128.log | awk 'NR%7==1' | tail -3
MSMB200_UMMA_DEBUG=1 timeout -k 5 300 python tools/k1_experiments.py --frames 8000000 kmajor:MSMB200_UMMA_MN=0 mn4:MSMB200_UMMA_MN_STAGES=4 mn5: > $O/r2ac_k1.log 2>&1
grep -v "^\[umma" $O/r2ac_k1.log | tail -3; grep "umma v2 dbg" $O/r2ac_k1.log | awk 'NR%7==1' | tail -3
timeout 1500 python bench.py > $O/r2ac_bench_1gpu.json 2> $O/r2ac_bench_1gpu.err
python - <<'PY'
import json
try:
    l=[x for x in open("gpurun_out/r2ac_bench_1gpu.json").read().splitlines() if x.startswith("{")][-1]
    d=json.loads(l); print("value %.1f M  step %.2f ms" % (d["value"]/1e6, d["ms_per_step"]), json.dumps(d["phases_ms"]), d["clocks"]["sm_mhz"])
    print("e2e %.1f M" % (d["e2e"]["value"]/1e6), "config2 %.3f ms" % d["other_configs"]["config2_tica_10Mx64"]["ms"], "config3 %.3f ms" % d["other_configs"]["config3_assign_10Mx16_k500"]["ms"])
except Exception as e:
    print("bench failed", e)
PY
