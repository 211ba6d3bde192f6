#!/bin/bash
# round 2, call E: swizzle / instruction probes, config 5 on one GPU
mkdir -p gpurun_out
O=gpurun_out
timeout -k 5 120 tools/_bin/umma_probe5 > $O/r2e_probe5.log 2>&1; echo "probe exit $?" >> $O/r2e_probe5.log
grep -i "swizzle\|one accumulator" $O/r2e_probe5.log
timeout -k 5 120 tools/_bin/umma_probe6 > $O/r2e_probe6.log 2>&1; echo "probe exit $?" >> $O/r2e_probe6.log
cat $O/r2e_probe6.log
timeout -k 5 600 python tools/config5_rmsd.py --frames 500000 --k 200 --templates 200 --check-k 200 > $O/r2e_config5_small.json 2> $O/r2e_config5_small.err; tail -2 $O/r2e_config5_small.err; cat $O/r2e_config5_small.json
timeout -k 5 900 python tools/config5_rmsd.py --check-k 100 > $O/r2e_config5_1gpu.json 2> $O/r2e_config5_1gpu.err; tail -2 $O/r2e_config5_1gpu.err; cat $O/r2e_config5_1gpu.json
