#!/bin/bash
# round 2, call I: launch list of a bench step + full captures of the K1 v2 kernel and the RMSD pass
mkdir -p gpurun_out
O=gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2i_launches_step.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-f64-check --no-ref-schedule > $O/r2i_ncu_launches.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(l for l in open("gpurun_out/r2i_launches_step.csv") if not l.startswith("==")))
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
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tica_umma_v2_kernel --launch-skip 1 --launch-count 1 \
   -o $O/r2i_k1_v2_full -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-f64-check --no-ref-schedule > $O/r2i_ncu_k1.log 2>&1
ls -la $O/*.ncu-rep | tail -3
