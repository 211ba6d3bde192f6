#!/bin/bash
# round 2, 2 GPUs: multi-GPU tests again after the per-device attribute fix
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests/test_gpu_devices.py tests/test_gpu_parallel.py -q > gpurun_out/r2p_pytest_multi.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2p_pytest_multi.log
tail -5 gpurun_out/r2p_pytest_multi.log
