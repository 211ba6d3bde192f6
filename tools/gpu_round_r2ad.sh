#!/bin/bash
# round 2, call AD: wide refine kernel (parity, timing), then the final artefacts of the round
mkdir -p gpurun_out
O=gpurun_out
timeout -k 5 900 python -m pytest tests -m gpu -q -x > $O/r2g_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/r2g_pytest_gpu.log
tail -3 $O/r2g_pytest_gpu.log
if ! grep -q "pytest exit 0" $O/r2g_pytest_gpu.log; then echo "GPU SUITE FAILED"; grep -E "^E |Error|assert|FAILED" $O/r2g_pytest_gpu.log | head -30; fi
timeout -k 5 400 python tools/assign_points.py > $O/r2g_assign_points.log 2>&1; echo "exit $?" >> $O/r2g_assign_points.log
cat $O/r2g_assign_points.log
MSMB200_ASSIGN_REFINE_FULL=1 timeout -k 5 200 python tools/assign_points.py 10000000,128,2000 2>&1 | tail -1
timeout -k 5 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2g_smoke.log 2>&1; tail -1 $O/r2g_smoke.log
timeout 1500 python bench.py > $O/r2g_bench_1gpu.json 2> $O/r2g_bench_1gpu.err
python - <<'PY'
import json
try:
    l=[x for x in open("gpurun_out/r2g_bench_1gpu.json").read().splitlines() if x.startswith("{")][-1]
    d=json.loads(l); print("value %.1f M  step %.2f ms" % (d["value"]/1e6, d["ms_per_step"]), json.dumps(d["phases_ms"]), d["clocks"]["sm_mhz"])
    print("e2e %.1f M" % (d["e2e"]["value"]/1e6), {k: (v.get("ms") or v.get("seconds")) for k, v in d["other_configs"].items()}, d["check"]["eig_err_vs_f64"])
except Exception as e:
    print("bench failed", e)
PY
Q="--no-cpu-baseline --no-e2e --no-f64-check --no-ref-schedule --no-other-configs"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tica_umma_v2_kernel --launch-skip 1 --launch-count 1 \
   -o $O/r2g_k1_mn_full -f python bench.py --steps 1 --warmup 1 $Q > $O/r2g_ncu_k1.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:tica_|kcenters_|candidate_|rmsd_|assign_' -c 400 --csv --log-file $O/r2g_launches_step.csv \
    python bench.py --steps 2 --warmup 1 $Q > $O/r2g_ncu_launches.log 2>&1
timeout -k 5 300 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:assign_' -c 40 --csv --log-file $O/r2g_launches_assign_stream.csv \
    python tools/profile_assign.py 10000000 128 2000 > $O/r2g_assign_times.log 2>&1
ls -la $O/*.ncu-rep | tail -3
