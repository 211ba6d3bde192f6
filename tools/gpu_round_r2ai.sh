#!/bin/bash
# round 2, call AI: MN-major ring of 4 vs 5 tiles in bench conditions (alternating, 3 runs each)
mkdir -p gpurun_out
O=gpurun_out
Q="--no-cpu-baseline --no-e2e --no-f64-check --no-ref-schedule --no-other-configs --steps 6 --warmup 3"
for i in 1 2 3; do
  MSMB200_UMMA_MN_STAGES=4 timeout 600 python bench.py $Q > $O/r2m_bench_mn4_$i.json 2> /dev/null
  MSMB200_UMMA_MN_STAGES=5 timeout 600 python bench.py $Q > $O/r2m_bench_mn5_$i.json 2> /dev/null
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2m_bench_mn*.json")):
    try:
        l=[x for x in open(f).read().splitlines() if x.startswith("{")][-1]
        d=json.loads(l); print(f.split("/")[-1], "value %.1f M  tica %.2f ms" % (d["value"]/1e6, d["phases_ms"]["tica_fit"]), d["phases_ms"]["tica_fit_steps"], d["clocks"]["sm_mhz"])
    except Exception as e:
        print(f, "failed", e)
PY
