#!/bin/bash
# round 2, call AB: K1 MN-major mode as the default, ring of 4 vs 5 tiles; full GPU suite; quick bench
mkdir -p gpurun_out
O=gpurun_out
MSMB200_UMMA_DEBUG=1 timeout -k 5 300 python tools/k1_experiments.py --frames 8000000 kmajor:MSMB200_UMMA_MN=0 mn4:MSMB200_UMMA_MN_STAGES=4 mn5:MSMB200_UMMA_MN_STAGES=5 > $O/r2ab_k1.log 2>&1
grep -v "^\[umma" $O/r2ab_k1.log | tail -4; grep "umma" $O/r2ab_k1.log | awk 'NR%7==1' | tail -3
timeout -k 5 900 python -m pytest tests -m gpu -q -x > $O/r2ab_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/r2ab_pytest_gpu.log
tail -3 $O/r2ab_pytest_gpu.log
B="--steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-other-configs --no-ref-schedule"
timeout 600 python bench.py $B > $O/r2ab_bench.json 2> $O/r2ab_bench.err
python - <<'PY'
import json
try:
    l=[x for x in open("gpurun_out/r2ab_bench.json").read().splitlines() if x.startswith("{")][-1]
    d=json.loads(l); print("value %.1f M  step %.2f ms" % (d["value"]/1e6, d["ms_per_step"]), json.dumps(d["phases_ms"]), d["clocks"]["sm_mhz"], json.dumps(d["check"])[:400])
except Exception as e:
    print("bench failed", e)
PY
