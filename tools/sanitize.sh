#!/bin/bash
# compute-sanitizer over the GPU parity tests (run on the GPU box: gpurun -- 'bash tools/sanitize.sh').
# memcheck on everything small enough; racecheck on the kernels that hand data between warps through
# shared memory (tcgen05 tICA converters, fused k-centers pass).  Logs -> gpurun_out/sanitizer_*.log
mkdir -p gpurun_out
S=/usr/local/cuda/bin/compute-sanitizer
timeout 1500 $S --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_lookahead.py -m gpu -x -q \
    -k "tiny or sqeuclidean or estimator" > gpurun_out/sanitizer_lookahead.log 2>&1; echo "memcheck lookahead: $?"
timeout 1500 $S --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_cluster.py tests/test_gpu_libdistance.py -m gpu -x -q \
    > gpurun_out/sanitizer_cluster.log 2>&1; echo "memcheck cluster/libdistance: $?"
timeout 1500 $S --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_tica.py -m gpu -x -q \
    -k "golden or ragged or f16 or reproducible" > gpurun_out/sanitizer_tica.log 2>&1; echo "memcheck tica: $?"
timeout 1500 $S --tool racecheck --error-exitcode 1 python -m pytest tests/test_gpu_lookahead.py -m gpu -x -q -k "tiny" \
    > gpurun_out/sanitizer_race_lookahead.log 2>&1; echo "racecheck lookahead: $?"
tail -3 gpurun_out/sanitizer_*.log
