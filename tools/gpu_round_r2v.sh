#!/bin/bash
# round 2, call V: look-ahead with three values per lane (two entries), T = 1024: parity, diagnostics, bench
mkdir -p gpurun_out
O=gpurun_out
timeout -k 5 400 python -m pytest tests/test_gpu_lookahead.py tests/test_gpu_cluster.py tests/test_gpu_widen.py -q -x > $O/r2v_pytest.log 2>&1; echo "pytest exit $?" >> $O/r2v_pytest.log
tail -4 $O/r2v_pytest.log
if ! grep -q "pytest exit 0" $O/r2v_pytest.log; then echo "PARITY FAILED / HANG, stopping"; grep -E "^E |Error|assert" $O/r2v_pytest.log | head -30; exit 1; fi
timeout -k 5 600 python tools/lookahead_diag.py --caps 512,1024,2048 > $O/r2v_lookahead_diag.log 2>&1; echo "exit $?" >> $O/r2v_lookahead_diag.log
head -12 $O/r2v_lookahead_diag.log
B="--steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-f64-check --no-other-configs"
timeout 600 python bench.py $B > $O/r2v_bench.json 2> $O/r2v_bench.err
python - <<'PY'
import json
try:
    l=[x for x in open("gpurun_out/r2v_bench.json").read().splitlines() if x.startswith("{")][-1]
    d=json.loads(l); print("value %.1f M  step %.2f ms" % (d["value"]/1e6, d["ms_per_step"]), json.dumps(d["phases_ms"]), d["clocks"]["sm_mhz"], json.dumps(d["value_reference_schedule"])[:200])
except Exception as e:
    print("bench failed", e)
PY
