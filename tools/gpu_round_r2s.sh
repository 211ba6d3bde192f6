#!/bin/bash
# round 2, call S: second-generation fused k-centers pass (kcenters_fused_pass_kernel): parity, then A/B against v1
mkdir -p gpurun_out
O=gpurun_out
timeout -k 5 600 python -m pytest tests/test_gpu_lookahead.py tests/test_gpu_cluster.py -q -x > $O/r2s_pytest.log 2>&1; echo "pytest exit $?" >> $O/r2s_pytest.log
tail -3 $O/r2s_pytest.log
if ! grep -q "pytest exit 0" $O/r2s_pytest.log; then echo "PARITY FAILED, stopping"; tail -40 $O/r2s_pytest.log; exit 1; fi
B="--steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-f64-check --no-ref-schedule --no-other-configs"
timeout 600 python bench.py $B > $O/r2s_bench_v2.json 2> $O/r2s_bench_v2.err
MSMB200_K2B_V1=1 timeout 600 python bench.py $B > $O/r2s_bench_v1.json 2> $O/r2s_bench_v1.err
python - <<'PY'
import json
for v in ("v2", "v1"):
    try:
        l=[x for x in open("gpurun_out/r2s_bench_%s.json" % v).read().splitlines() if x.startswith("{")][-1]
        d=json.loads(l); print(v, "value %.1f M  step %.2f ms" % (d["value"]/1e6, d["ms_per_step"]), json.dumps(d["phases_ms"]), d["clocks"]["sm_mhz"])
    except Exception as e:
        print(v, "failed", e)
PY
