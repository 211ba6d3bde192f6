#!/bin/bash
# round 2, call AF: one pinned arena per device (no cudaMallocHost inside the timed steps), MN-major mode at D = 64
mkdir -p gpurun_out
O=gpurun_out
timeout -k 5 900 python -m pytest tests -m gpu -q -x > $O/r2i_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/r2i_pytest_gpu.log
tail -3 $O/r2i_pytest_gpu.log
if ! grep -q "pytest exit 0" $O/r2i_pytest_gpu.log; then grep -E "^E |Error|assert|FAILED" $O/r2i_pytest_gpu.log | head -30; fi
MSMB200_UMMA_DEBUG=1 timeout -k 5 300 python tools/k1_experiments.py --frames 10000000 --features 64 kmajor:MSMB200_UMMA_MN64=0 mn5: > $O/r2i_k1_d64.log 2>&1
grep -v "^\[umma" $O/r2i_k1_d64.log | tail -2; grep "umma v2 dbg" $O/r2i_k1_d64.log | awk 'NR%7==1' | tail -2
Q="--no-cpu-baseline --no-e2e --no-f64-check --no-ref-schedule --no-other-configs"
for i in 1 2; do timeout 600 python bench.py $Q > $O/r2i_bench_quick_$i.json 2> /dev/null; done
timeout 1500 python bench.py > $O/r2i_bench_1gpu.json 2> $O/r2i_bench_1gpu.err
timeout -k 5 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2i_smoke.log 2>&1; tail -1 $O/r2i_smoke.log
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2i_bench_*.json")):
    try:
        l=[x for x in open(f).read().splitlines() if x.startswith("{")][-1]
        d=json.loads(l); print(f.split("/")[-1], "value %.1f M  step %.2f ms" % (d["value"]/1e6, d["ms_per_step"]), d["phases_ms"]["tica_fit_steps"], d["phases_ms"]["kcenters_fit_steps"], d["clocks"]["sm_mhz"])
        if d.get("other_configs"): print("   e2e %.1f M" % (d["e2e"]["value"]/1e6), {k: (v.get("ms") or v.get("seconds")) for k, v in d["other_configs"].items()}, d["check"]["eig_err_vs_f64"])
    except Exception as e:
        print(f, "failed", e)
PY
