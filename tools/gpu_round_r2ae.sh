#!/bin/bash
# round 2, call AE: the lead-in step of bench.py (first-step outliers), final bench line
mkdir -p gpurun_out
O=gpurun_out
Q="--no-cpu-baseline --no-e2e --no-f64-check --no-ref-schedule --no-other-configs"
for i in 1 2; do
  timeout 600 python bench.py $Q --no-lead-in > $O/r2h_bench_nolead_$i.json 2> /dev/null
  timeout 600 python bench.py $Q > $O/r2h_bench_lead_$i.json 2> /dev/null
done
timeout 1500 python bench.py > $O/r2h_bench_1gpu.json 2> $O/r2h_bench_1gpu.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2h_bench_*.json")):
    try:
        l=[x for x in open(f).read().splitlines() if x.startswith("{")][-1]
        d=json.loads(l); print(f.split("/")[-1], "value %.1f M  step %.2f ms" % (d["value"]/1e6, d["ms_per_step"]), d["phases_ms"]["tica_fit_steps"], d["phases_ms"]["kcenters_fit_steps"], d["clocks"])
    except Exception as e:
        print(f, "failed", e)
PY
