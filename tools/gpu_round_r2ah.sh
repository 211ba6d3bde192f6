#!/bin/bash
# round 2, call AH (2 GPUs): final tree -- NCCL / thread-communicator parity tests and the 2-GPU bench
mkdir -p gpurun_out
O=gpurun_out
timeout -k 5 900 python -m pytest tests/test_gpu_parallel.py tests/test_gpu_devices.py -m gpu -x -q > $O/r2l_pytest_multi_2gpu.log 2>&1; echo "pytest exit $?" >> $O/r2l_pytest_multi_2gpu.log
tail -3 $O/r2l_pytest_multi_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline --no-other-configs --no-e2e > $O/r2l_bench_2gpu.json 2> $O/r2l_bench_2gpu.err
python - <<'PY'
import json
try:
    l=[x for x in open("gpurun_out/r2l_bench_2gpu.json").read().splitlines() if x.startswith("{")][-1]
    d=json.loads(l); print("2 GPUs: value %.1f M  step %.2f ms" % (d["value"]/1e6, d["ms_per_step"]), json.dumps(d["phases_ms"]), d["check"]["eig_err_vs_f64"], d["check"]["kcenters_ids"], d["value_reference_schedule"]["ids_equal_lookahead"])
except Exception as e:
    print("bench failed", e)
PY
tail -2 $O/r2l_bench_2gpu.err
