#!/bin/bash
# round 2, call Q: deeper raw ring at D = 64 / 128, 8 upload threads
mkdir -p gpurun_out
O=gpurun_out
timeout -k 5 600 python -m pytest tests/test_gpu_tica.py tests/test_gpu_widen.py -q -x > $O/r2q_pytest.log 2>&1; echo "pytest exit $?" >> $O/r2q_pytest.log
tail -3 $O/r2q_pytest.log
if grep -q "pytest exit 124\|pytest exit 137" $O/r2q_pytest.log; then echo "HANG"; exit 1; fi
MSMB200_UMMA_DEBUG=1 timeout -k 5 600 python tools/k1_experiments.py --frames 10000000 --features 64 v2: > $O/r2q_k1_d64.log 2>&1
grep -v "^\[umma" $O/r2q_k1_d64.log | tail -1; grep "umma v2 dbg" $O/r2q_k1_d64.log | tail -1
MSMB200_UMMA_DEBUG=1 timeout -k 5 600 python tools/k1_experiments.py --frames 10000000 --features 32 --engine umma_3xf16 v2: > $O/r2q_k1_d32.log 2>&1
grep -v "^\[umma" $O/r2q_k1_d32.log | tail -1; grep "umma v2 dbg" $O/r2q_k1_d32.log | tail -1
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-f64-check --no-ref-schedule --no-other-configs > $O/r2q_bench_e2e.json 2> $O/r2q_bench_e2e.err
python - <<'PY'
import json
l=[x for x in open("gpurun_out/r2q_bench_e2e.json").read().splitlines() if x.startswith("{")][-1]
d=json.loads(l); print("value", d["value"], "e2e", json.dumps(d["e2e"])[:500])
PY
