#!/bin/bash
# round 2, call AJ: edge kernel with 4 rows per barrier pair -- GPU suite, quick bench, launch list of the tICA phase
mkdir -p gpurun_out
O=gpurun_out
timeout -k 5 900 python -m pytest tests -m gpu -q > $O/r2n_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/r2n_pytest_gpu.log
tail -3 $O/r2n_pytest_gpu.log
if ! grep -q "pytest exit 0" $O/r2n_pytest_gpu.log; then grep -E "^E |Error|assert|FAILED" $O/r2n_pytest_gpu.log | head -30; fi
timeout -k 5 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2n_smoke.log 2>&1; tail -1 $O/r2n_smoke.log
Q="--no-cpu-baseline --no-e2e --no-ref-schedule --no-other-configs --steps 6 --warmup 3"
for i in 1 2; do timeout 600 python bench.py $Q > $O/r2n_bench_quick_$i.json 2> /dev/null; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:tica_' -c 60 --csv --log-file $O/r2n_launches_tica.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-f64-check --no-ref-schedule --no-other-configs > /dev/null 2>&1
python - <<'PY'
import json, glob, csv
for f in sorted(glob.glob("gpurun_out/r2n_bench_quick_*.json")):
    try:
        l=[x for x in open(f).read().splitlines() if x.startswith("{")][-1]
        d=json.loads(l); print(f.split("/")[-1], "value %.1f M  tica %.2f ms kc %.2f" % (d["value"]/1e6, d["phases_ms"]["tica_fit"], d["phases_ms"]["kcenters_fit"]), d["phases_ms"]["tica_fit_steps"], d["check"]["eig_err_vs_f64"], d["clocks"]["sm_mhz"])
    except Exception as e:
        print(f, "failed", e)
rows=[r for r in csv.reader(open("gpurun_out/r2n_launches_tica.csv")) if len(r)>5]
h=[i for i,r in enumerate(rows) if r[0]=="ID"][0]; ki=rows[h].index("Kernel Name"); vi=rows[h].index("Metric Value")
for r in rows[h+2:]:
    if "edges" in r[ki] or "v2_kernel" in r[ki]: print(r[ki][:50], r[vi])
PY
