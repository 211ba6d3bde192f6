#!/bin/bash
# compute-sanitizer over the kernels of the last session of round 2 (run on the GPU box): memcheck on the MN-major
# K1 mode (golden / ragged / rescue / reproducible tests run it at D = 128 and 256 with lag <= 32), on the look-ahead
# passes (first pass, second-generation fused pass, three-value lane records in shared memory, select / chain), on both
# tensor-core assign filters (resident and streamed centres) and the wide re-scan kernel; racecheck on the look-ahead
# passes and the streamed assign filter.  Logs -> gpurun_out/r2j_sanitizer_*.log
mkdir -p gpurun_out
S=/usr/local/cuda/bin/compute-sanitizer
run() {   # name, tool, pytest args...
    local name=$1 tool=$2; shift 2
    timeout -k 10 500 $S --tool $tool --error-exitcode 1 python -m pytest "$@" -m gpu -x -q -p no:cacheprovider \
        > gpurun_out/r2j_sanitizer_${tool}_${name}.log 2>&1
    echo "$tool $name: exit $?" | tee -a gpurun_out/r2j_sanitizer_summary.log
    grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r2j_sanitizer_${tool}_${name}.log | tail -2 | tee -a gpurun_out/r2j_sanitizer_summary.log
}
: > gpurun_out/r2j_sanitizer_summary.log
run tica_mn memcheck tests/test_gpu_tica.py -k "golden or ragged or rescue or reproducible or f16_engine_matches"
run lookahead memcheck tests/test_gpu_lookahead.py -k "tiny or sqeuclidean or estimator or equals_pass_per_centre"
run assign memcheck tests/test_gpu_libdistance.py -k "tensor_core or streamed"
run lookahead racecheck tests/test_gpu_lookahead.py -k "tiny or sqeuclidean"
run assign_stream racecheck tests/test_gpu_libdistance.py -k "streamed_centres_filter_vs_reference or streamed_kernel_on"
