#!/bin/bash
# round 2, call L: RMSD tile warps with TMA bulk copies, K1 idle waits
mkdir -p gpurun_out
O=gpurun_out
timeout -k 5 900 python -m pytest tests/test_gpu_rmsd.py tests/test_gpu_tica.py -q -x > $O/r2l_pytest.log 2>&1; echo "pytest exit $?" >> $O/r2l_pytest.log
tail -4 $O/r2l_pytest.log
if grep -q "pytest exit 124\|pytest exit 137" $O/r2l_pytest.log; then echo "HANG"; exit 1; fi
timeout -k 5 900 python tools/config5_rmsd.py --check-k 100 > $O/r2l_config5_1gpu.json 2> $O/r2l_config5_1gpu.err; tail -2 $O/r2l_config5_1gpu.err; cat $O/r2l_config5_1gpu.json
MSMB200_UMMA_DEBUG=1 timeout -k 5 600 python tools/k1_experiments.py --frames 8000000 v2: v1:MSMB200_UMMA_V1=1 > $O/r2l_k1_experiments.log 2>&1
grep -v "^\[umma" $O/r2l_k1_experiments.log | tail -3; grep "umma" $O/r2l_k1_experiments.log | awk 'NR%7==1' | tail -2
MSMB200_UMMA_DEBUG=1 timeout -k 5 600 python tools/k1_experiments.py --frames 10000000 --features 64 v2: > $O/r2l_k1_experiments_d64.log 2>&1
grep -v "^\[umma" $O/r2l_k1_experiments_d64.log | tail -1; grep "umma v2 dbg" $O/r2l_k1_experiments_d64.log | tail -1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:tica_|kcenters_|candidate_|rmsd_' -c 400 --csv --log-file $O/r2l_launches_step.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-f64-check --no-ref-schedule > $O/r2l_ncu_launches.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(l for l in open("gpurun_out/r2l_launches_step.csv") if not l.startswith("==")))
hdr = rows[0]
ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    if len(r) <= iv: continue
    k = r[ik][:60]
    v = float(r[iv].replace(",", ""))
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(v for _, v in agg.values())
for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
    print("%-62s %4d launches %10.3f ms total %6.1f %%" % (k, n, v / 1e6, 100 * v / tot))
PY
