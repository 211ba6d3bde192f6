#!/bin/bash
# round 2, 8 GPUs of one box: multi-GPU tests (NCCL and in-process), config 5 at size, bench at 8
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L | head -8 > $O/r2m_gpus.log
timeout -k 5 900 python -m pytest tests/test_gpu_devices.py tests/test_gpu_parallel.py -q > $O/r2m_pytest_multi.log 2>&1; echo "pytest exit $?" >> $O/r2m_pytest_multi.log
tail -6 $O/r2m_pytest_multi.log
timeout -k 5 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 \
    tools/config5_rmsd.py --check-k 100 > $O/r2m_config5_8gpu.json 2> $O/r2m_config5_8gpu.err
tail -2 $O/r2m_config5_8gpu.err; cat $O/r2m_config5_8gpu.json
timeout -k 5 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 \
    bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline > $O/r2m_bench_8gpu.json 2> $O/r2m_bench_8gpu.err
tail -c 900 $O/r2m_bench_8gpu.json; tail -3 $O/r2m_bench_8gpu.err
timeout -k 5 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 \
    bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > $O/r2m_bench_2gpu.json 2> $O/r2m_bench_2gpu.err
tail -c 400 $O/r2m_bench_2gpu.json
timeout -k 5 600 python tools/devices_e2e.py > $O/r2m_devices_e2e.json 2> $O/r2m_devices_e2e.err; tail -3 $O/r2m_devices_e2e.err; cat $O/r2m_devices_e2e.json
