#!/bin/bash
# compute-sanitizer over the kernels that carry the round-2 numbers (run on the GPU box).  memcheck on the
# tcgen05 tICA engine (v2, both CTA-group variants, float64 rescue), the RMSD tile / solver / pruned passes,
# the tensor-core assign filter and the look-ahead k-centers; racecheck on the kernels that hand data
# between warps through shared memory.  Logs -> gpurun_out/r2s_sanitizer_*.log
mkdir -p gpurun_out
S=/usr/local/cuda/bin/compute-sanitizer
run() {   # name, tool, pytest args...
    local name=$1 tool=$2; shift 2
    timeout -k 10 700 $S --tool $tool --error-exitcode 1 python -m pytest "$@" -m gpu -x -q -p no:cacheprovider \
        > gpurun_out/r2s_sanitizer_${tool}_${name}.log 2>&1
    echo "$tool $name: exit $?" | tee -a gpurun_out/r2s_sanitizer_summary.log
    grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r2s_sanitizer_${tool}_${name}.log | tail -2 | tee -a gpurun_out/r2s_sanitizer_summary.log
}
: > gpurun_out/r2s_sanitizer_summary.log
run tica memcheck tests/test_gpu_tica.py -k "golden or ragged or rescue or narrow or reproducible"
run rmsd memcheck tests/test_gpu_rmsd.py
run assign memcheck tests/test_gpu_libdistance.py -k "tensor_core or assign_nearest"
run lookahead memcheck tests/test_gpu_lookahead.py -k "tiny or sqeuclidean or estimator"
run rmsd racecheck tests/test_gpu_rmsd.py -k "bit_for_bit or pruned"
run lookahead racecheck tests/test_gpu_lookahead.py -k "tiny"
run tica racecheck tests/test_gpu_tica.py -k "narrow"
