#!/bin/bash
# round 2, call X: lane-owns-frames fused pass, two frames per lane, swizzled 32-float slices
mkdir -p gpurun_out
O=gpurun_out
timeout -k 5 400 python -m pytest tests/test_gpu_lookahead.py tests/test_gpu_cluster.py tests/test_gpu_widen.py -q -x > $O/r2x_pytest.log 2>&1; echo "pytest exit $?" >> $O/r2x_pytest.log
tail -4 $O/r2x_pytest.log
if ! grep -q "pytest exit 0" $O/r2x_pytest.log; then echo "PARITY FAILED / HANG, stopping"; grep -E "^E |Error|assert" $O/r2x_pytest.log | head -30; exit 1; fi
timeout -k 5 600 python tools/lookahead_diag.py --caps 1024 > $O/r2x_lookahead_diag.log 2>&1; echo "exit $?" >> $O/r2x_lookahead_diag.log
head -3 $O/r2x_lookahead_diag.log
MSMB200_K2B_COOP=1 timeout -k 5 600 python tools/lookahead_diag.py --caps 1024 2>&1 | head -1
B="--steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-f64-check --no-other-configs --no-ref-schedule"
timeout 600 python bench.py $B > $O/r2x_bench.json 2> $O/r2x_bench.err
MSMB200_K2B_COOP=1 timeout 600 python bench.py $B > $O/r2x_bench_coop.json 2> $O/r2x_bench_coop.err
python - <<'PY'
import json
for v in ("r2x_bench", "r2x_bench_coop"):
    try:
        l=[x for x in open("gpurun_out/%s.json" % v).read().splitlines() if x.startswith("{")][-1]
        d=json.loads(l); print(v, "value %.1f M  step %.2f ms" % (d["value"]/1e6, d["ms_per_step"]), json.dumps(d["phases_ms"]), d["clocks"]["sm_mhz"])
    except Exception as e:
        print(v, "failed", e)
PY
