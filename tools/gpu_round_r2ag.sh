#!/bin/bash
# round 2, call AG: final tree -- GPU suite, smoke, look-ahead pass times, bench line
mkdir -p gpurun_out
O=gpurun_out
timeout -k 5 900 python -m pytest tests -m gpu -q > $O/r2k_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/r2k_pytest_gpu.log
tail -3 $O/r2k_pytest_gpu.log
if ! grep -q "pytest exit 0" $O/r2k_pytest_gpu.log; then grep -E "^E |Error|assert|FAILED" $O/r2k_pytest_gpu.log | head -30; fi
timeout -k 5 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2k_smoke.log 2>&1; tail -1 $O/r2k_smoke.log
timeout -k 5 300 python tools/lookahead_diag.py --caps 1024 2>&1 | head -1
timeout 1500 python bench.py > $O/r2k_bench_1gpu.json 2> $O/r2k_bench_1gpu.err
python - <<'PY'
import json
try:
    l=[x for x in open("gpurun_out/r2k_bench_1gpu.json").read().splitlines() if x.startswith("{")][-1]
    d=json.loads(l); print("value %.1f M  step %.2f ms" % (d["value"]/1e6, d["ms_per_step"]), json.dumps(d["phases_ms"]), d["clocks"]["sm_mhz"])
    print("   e2e %.1f M" % (d["e2e"]["value"]/1e6), {k: (v.get("ms") or v.get("seconds")) for k, v in d["other_configs"].items()}, d["check"]["eig_err_vs_f64"])
except Exception as e:
    print("bench failed", e)
PY
